"""Parity of the CUDA path with the oracle and the reference's golden vectors.  Everything here goes
through the C ABI (include/eigkl.h via eig_kl_algorithm_b200.api) on a real B200.

Bars: integer / fp32-ordered work bit-exact (KL graph layout and weights, D-values, initial cut,
gain column, swap sequence, trace file bytes); fp64 Fiedler solve within the north-star tolerance
(lambda2 rel. err <= 1e-8, sine <= 1e-6 after sign alignment).
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, CIRCUITS
from eig_kl_algorithm_b200 import api, datasets
from test_oracle import APPENDIX_D, golden_swaps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handles(eigkl_lib, circuits):
    hs = {}
    for c in CIRCUITS:
        h = api.Handle()
        h.load_hgr(circuits[c])
        hs[c] = h
    yield hs
    for h in hs.values():
        h.close()


# ---------------------------------------------------------------------------------------------------
# assembly
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c", CIRCUITS)
def test_kl_graph_layout_bit_exact(c, handles, oracle, circuits):
    h = handles[c]
    h.assemble_kl_graph()
    rp, fe, col, w = h.get_kl_graph()
    o = oracle.OracleKL(oracle.OracleHgr(circuits[c]))
    assert np.array_equal(rp, o.rowptr.astype(np.int32))
    assert np.array_equal(fe, o.fwd_end.astype(np.int32))
    assert np.array_equal(col, o.col)                       # traversal order incl. unordered_map order
    assert np.array_equal(w.view(np.uint32), o.w.view(np.uint32))   # fp32 sums in file order, bit for bit


def _laplacian_in_file_ids(h):
    """The assembled matrix is stored in the EIG stage's node order; bring it back to the file's ids."""
    import scipy.sparse as sp
    rp, col, val = h.get_laplacian()
    perm = h.node_order()
    n = h.n_nodes
    rows = np.repeat(np.arange(n), np.diff(rp))
    a = sp.coo_matrix((val, (perm[rows], perm[col])), shape=(n, n)).tocsr()
    a.sort_indices()
    return a, perm


@pytest.mark.parametrize("c", CIRCUITS)
@pytest.mark.parametrize("natural", [False, True])
def test_laplacian_matches_oracle(c, natural, oracle, circuits):
    o = oracle.OracleEIG(oracle.OracleHgr(circuits[c]))
    with api.Handle(flags=api.EIGKL_F_NATURAL_ORDER if natural else 0) as h:
        h.load_hgr(circuits[c])
        h.assemble_laplacian()
        a, perm = _laplacian_in_file_ids(h)
        assert sorted(perm.tolist()) == list(range(h.n_nodes))
        assert natural == bool(np.array_equal(perm, np.arange(h.n_nodes)))
        assert np.array_equal(a.indptr, o.rowptr.astype(a.indptr.dtype))
        assert np.array_equal(a.indices, o.col)
        rows = np.repeat(np.arange(h.n_nodes), np.diff(a.indptr))
        offd = rows != a.indices
        assert np.array_equal(a.data[offd], o.val[offd])        # same fp64 sums in the same (file) order
        # the diagonal is -(row sum) taken in the stored column order, which the renumbering changes
        assert np.allclose(a.data[~offd], o.val[~offd], rtol=1e-12, atol=0)      # up-to-900-term row sums, reordered
        if natural:
            assert np.array_equal(a.data, o.val)
        x = np.random.default_rng(0).standard_normal(h.n_nodes)
        y = h.spmv(x)                                            # takes and returns vectors in file ids
        yo = o.spmv(x)
        assert np.abs(y - yo).max() <= 1e-12 * np.abs(yo).max()
        assert np.abs(h.spmv(np.ones(h.n_nodes))).max() < 1e-13 * np.abs(o.val).max() * 64   # L 1 = 0 up to row round-off


# ---------------------------------------------------------------------------------------------------
# KL
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c", CIRCUITS)
def test_dvalues_and_cut_bit_exact(c, handles, oracle, circuits, workdir):
    h = handles[c]
    h.assemble_kl_graph()
    o = oracle.OracleKL(oracle.OracleHgr(circuits[c]))
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), h.n_nodes)
    rng = np.random.default_rng(5)
    for side in (g["side"], rng.integers(0, 2, h.n_nodes).astype(np.uint8)):
        h.set_partition(side)
        assert np.array_equal(h.dvalues().view(np.uint32), o.dvalues(side).view(np.uint32))
        assert h.cut().view(np.uint32) == o.cut0(side).view(np.uint32)


@pytest.mark.parametrize("c", CIRCUITS)
def test_kl_trace_byte_exact_vs_reference(c, handles, workdir, tmp_path):
    """cKL <c>.hgr -EIG: the trace file equals the file the reference wrote on one core, byte for byte,
    and the swapped node ids equal the instrumented reference's."""
    h = handles[c]
    h.assemble_kl_graph()
    h.load_eig(datasets.golden_eig_path(workdir, c))
    tr = h.kl_run()
    out = str(tmp_path / "trace.txt")
    api.write_trace(out, tr)
    raw = open(out, "rb").read()
    assert tr["swaps"] == APPENDIX_D[c][0]
    n1, n2 = golden_swaps(c)
    assert np.array_equal(tr["node1"][1:], n1) and np.array_equal(tr["node2"][1:], n2)
    assert raw == open(os.path.join(GOLDEN, c + ".kl_trace_1core.txt"), "rb").read()
    assert hashlib.md5(raw).hexdigest() == APPENDIX_D[c][1]
    side = h.get_partition()
    assert int(side.sum()) == h.n_nodes - (h.n_nodes + 1) // 2 or int(side.sum()) > 0


@pytest.mark.parametrize("c", CIRCUITS)
@pytest.mark.parametrize("variant", ["shared_bits", "global_bits", "warp_shared_bits", "warp_global_bits", "global_state"])
def test_kl_loop_variants_byte_exact(c, variant, circuits, workdir, tmp_path, monkeypatch):
    """The single-GPU swap loops -- tile keys + side bits in shared memory (default up to 524 288 nodes), tile keys in
    shared memory with the state bytes in global memory (up to 2 M nodes), each in its flat (default: a lane per
    neighbour row) and warp-per-row form, and everything in global memory (the cluster kernel) -- all reproduce the
    reference's one-core trace byte for byte."""
    if variant.endswith("global_bits"):
        monkeypatch.setenv("EIGKL_KL_GBITS", "1")
    if variant.startswith("warp_"):
        monkeypatch.setenv("EIGKL_KL_FLAT", "0")
    if variant == "global_state":
        monkeypatch.setenv("EIGKL_KL_LOCAL", "0")
    with api.Handle() as h:
        h.load_hgr(circuits[c])
        h.assemble_kl_graph()
        h.load_eig(datasets.golden_eig_path(workdir, c))
        tr = h.kl_run()
        st = h.stats()
        assert st["kl_local"] == {"shared_bits": 1, "global_bits": 2, "warp_shared_bits": 1, "warp_global_bits": 2, "global_state": 0}[variant]
        assert st["kl_flat"] == (1 if variant in ("shared_bits", "global_bits") else 0)
    out = str(tmp_path / "trace.txt")
    api.write_trace(out, tr)
    raw = open(out, "rb").read()
    assert raw == open(os.path.join(GOLDEN, c + ".kl_trace_1core.txt"), "rb").read()
    n1, n2 = golden_swaps(c)
    assert np.array_equal(tr["node1"][1:], n1) and np.array_equal(tr["node2"][1:], n2)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_kl_cluster_sizes_agree(cluster, oracle, circuits, workdir):
    c = "industry2"
    with api.Handle(kl_cluster=cluster) as h:
        h.load_hgr(circuits[c])
        h.assemble_kl_graph()
        h.load_eig(datasets.golden_eig_path(workdir, c))
        tr = h.kl_run()
        assert h.stats()["kl_cluster"] == cluster
    n1, n2 = golden_swaps(c)
    assert np.array_equal(tr["node1"][1:], n1) and np.array_equal(tr["node2"][1:], n2)


@pytest.mark.parametrize("c", ["fract", "ibm01"])
def test_kl_random_ordered_partition(c, handles, oracle, circuits):
    """The non -EIG branch (cKL.cpp:175-193): remain[] in shuffled order, ties go to the earlier position."""
    h = handles[c]
    h.assemble_kl_graph()
    o = oracle.OracleKL(oracle.OracleHgr(circuits[c]))
    rng = np.random.default_rng(11)
    perm = rng.permutation(h.n_nodes).astype(np.int32)
    mid = h.n_nodes // 2
    o0, o1 = perm[:mid], perm[mid:]
    side = np.zeros(h.n_nodes, np.uint8)
    side[o1] = 1
    ro = o.run(side, o0, o1)
    h.set_partition_ordered(o0, o1)
    tr = h.kl_run()
    assert tr["swaps"] == ro["swaps"]
    assert np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["node2"], ro["node2"])
    assert np.array_equal(tr["gain"].view(np.uint32), ro["gain"].view(np.uint32))
    assert np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))
    assert np.array_equal(h.get_partition(), ro["side"])


def _random_hypergraph(rng, n_nodes, n_nets, big=0):
    nets = []
    for _ in range(n_nets):
        k = int(rng.choice([1, 2, 2, 2, 3, 4, 5, 8]))
        nets.append(rng.choice(n_nodes, size=min(k, n_nodes), replace=False))
    for _ in range(big):
        nets.append(rng.choice(n_nodes, size=min(n_nodes, int(rng.integers(100, 400))), replace=False))
    nets.insert(len(nets) // 2, np.zeros(0, np.int64))       # an empty net line
    off = np.zeros(len(nets) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in nets])
    pins = np.concatenate(nets).astype(np.int32) if off[-1] else np.zeros(0, np.int32)
    return off, pins


@pytest.mark.parametrize("seed,n_nodes,n_nets,big", [(1, 50, 40, 0), (2, 700, 900, 3), (3, 5000, 5000, 6), (4, 300, 10, 2)])
def test_ragged_random_hypergraphs(seed, n_nodes, n_nets, big, oracle, tmp_path):
    """Ragged inputs: 1-pin and empty nets, isolated nodes, nets of hundreds of pins (rehash chains)."""
    rng = np.random.default_rng(seed)
    off, pins = _random_hypergraph(rng, n_nodes, n_nets, big)
    path = str(tmp_path / "r.hgr")
    with open(path, "w") as f:
        f.write(f"{len(off) - 1} {n_nodes}\n")
        for e in range(len(off) - 1):
            f.write(" ".join(str(int(p) + 1) for p in pins[off[e]:off[e + 1]]) + " \n")
    oh = oracle.OracleHgr(path)
    o = oracle.OracleKL(oh)
    side = rng.integers(0, 2, n_nodes).astype(np.uint8)
    ro = o.run(side)
    with api.Handle() as h:
        h.set_pins(n_nodes, off, pins)                       # the host-array entry point
        h.assemble_kl_graph()
        rp, fe, col, w = h.get_kl_graph()
        assert np.array_equal(rp, o.rowptr.astype(np.int32)) and np.array_equal(col, o.col)
        assert np.array_equal(w.view(np.uint32), o.w.view(np.uint32))
        h.set_partition(side)
        tr = h.kl_run()
        assert tr["swaps"] == ro["swaps"]
        assert np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["node2"], ro["node2"])
        assert np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))
        assert np.array_equal(tr["gain"].view(np.uint32), ro["gain"].view(np.uint32))


def test_long_run_of_isolated_nodes(oracle, tmp_path):
    """Thousands of consecutive nodes without a pin: their rows are empty, and a row block of the D-value kernel must not
    collect more of them than it can stage row offsets for (blocks are cut by non-zeros plus rows)."""
    rng = np.random.default_rng(11)
    n_nodes = 12000
    live = np.concatenate([np.arange(0, 1500), np.arange(9000, 12000)])      # nodes 1500..8999 never appear
    nets = [rng.choice(live, size=int(rng.choice([2, 2, 3, 4])), replace=False) for _ in range(6000)]
    off = np.zeros(len(nets) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in nets])
    pins = np.concatenate(nets).astype(np.int32)
    path = str(tmp_path / "iso.hgr")
    with open(path, "w") as f:
        f.write(f"{len(nets)} {n_nodes}\n")
        for x in nets:
            f.write(" ".join(str(int(q) + 1) for q in x) + "\n")
    oh = oracle.OracleHgr(path)
    o = oracle.OracleKL(oh)
    side = rng.integers(0, 2, n_nodes).astype(np.uint8)
    with api.Handle() as h:
        h.set_pins(n_nodes, off, pins)
        h.assemble_kl_graph()
        h.set_partition(side)
        assert np.array_equal(h.dvalues().view(np.uint32), o.dvalues(side).view(np.uint32))
        tr = h.kl_run()
    ro = o.run(side)
    assert tr["swaps"] == ro["swaps"] and np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["node2"], ro["node2"])
    assert np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))


def test_bad_inputs_are_rejected(tmp_path):
    with api.Handle() as h:
        off = np.array([0, 3], np.int64)
        with pytest.raises(api.EigklError) as e:
            h.set_pins(4, off, np.array([0, 1, 7], np.int32))            # pin out of range
        assert e.value.code == -3
        h.set_pins(4, off, np.array([0, 1, 1], np.int32))                # duplicate pin in a net
        with pytest.raises(api.EigklError) as e:
            h.assemble_kl_graph()
        assert e.value.code == -3
        with pytest.raises(api.EigklError) as e:
            h.load_hgr(str(tmp_path / "missing.hgr"))
        assert e.value.code == -2
        with pytest.raises(api.EigklError) as e:                        # offsets that go backwards: negative net sizes
            h.set_pins(4, np.array([0, 3, 2, 4], np.int64), np.array([0, 1, 2, 3], np.int32))
        assert e.value.code == -3
        h.set_pins(4, off, np.array([0, 1, 2], np.int32))
        with pytest.raises(api.EigklError):
            h.kl_run()                                                   # no graph / partition yet


def test_second_pass_without_new_partition_is_a_real_pass(handles, oracle, circuits, workdir):
    """KL() rebuilds remain[] / split[] from the current sides on every call (cKL.cpp:290-301): a second eigkl_kl_run
    without a new partition starts from the FINAL partition of the first, with fresh locks and ascending orders."""
    import ctypes as C
    c = "ibm01"
    h = handles[c]
    h.assemble_kl_graph()
    h.load_eig(datasets.golden_eig_path(workdir, c))
    tr1 = h.kl_run()
    side1 = h.get_partition()
    tr2 = h.kl_run()                                                     # no set_partition in between
    o = oracle.OracleKL(oracle.OracleHgr(circuits[c]))
    ro = o.run(side1)
    assert tr2["swaps"] == ro["swaps"] and np.array_equal(tr2["node1"], ro["node1"]) and np.array_equal(tr2["node2"], ro["node2"])
    assert np.array_equal(tr2["cut"].view(np.uint32), ro["cut"].view(np.uint32))
    assert np.array_equal(h.get_partition(), ro["side"])
    # cut / D-values after a pass describe the final partition too
    assert h.cut().view(np.uint32) == o.cut0(ro["side"]).view(np.uint32)
    assert np.array_equal(h.dvalues().view(np.uint32), o.dvalues(ro["side"]).view(np.uint32))
    # a trace that is too short is refused before the pass runs
    h.load_eig(datasets.golden_eig_path(workdir, c))
    cut = np.zeros(4, np.float32)
    t = api.Trace(4, 0, cut.ctypes.data_as(C.POINTER(C.c_float)), None, None, None)
    assert h.lib.eigkl_kl_run(h._h, C.byref(t)) == -1
    assert h.stats()["kl_swaps"] == tr2["swaps"]                         # nothing ran


# ---------------------------------------------------------------------------------------------------
# EIG
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c", ["fract", "ibm01", "industry2"])
def test_fiedler_vs_golden(c, handles, oracle, workdir, tmp_path):
    h = handles[c]
    h.assemble_laplacian()
    lam, v = h.fiedler()
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), h.n_nodes)
    st = h.stats()
    assert st["converged"] == 1
    assert abs(lam - g["lambda2"]) / g["lambda2"] <= 1e-8                 # north star: 1e-8 relative
    assert abs(np.linalg.norm(v) - 1.0) < 1e-12
    cs = abs(v @ g["vec"]) / np.linalg.norm(g["vec"])
    assert np.sqrt(max(0.0, 1.0 - cs * cs)) <= 1e-6                       # north star: 1e-6 sine
    # residual of the returned pair
    r = np.linalg.norm(h.spmv(v) - lam * v)
    assert r <= 1e-9 * max(1.0, abs(lam)) * 1e2
    # median / sides on the device == cEIG.cpp:55-65,218 on the host
    med, side = h.partition_from_fiedler()
    assert med == oracle.median(v)
    assert np.array_equal(side, (med > v).astype(np.uint8))
    s = np.sign(v @ g["vec"])
    side_aligned = side if s > 0 else (oracle.median(-v) > -v).astype(np.uint8)
    assert int((side_aligned != g["side"]).sum()) == 0                    # same partition as the reference's file
    # the file we write parses back to the same numbers (12 significant digits)
    out = str(tmp_path / "eig.txt")
    h.write_eig(out)
    back = oracle.read_eig(out, h.n_nodes)
    assert np.array_equal(back["side"], side)
    assert abs(back["lambda2"] - lam) <= 1e-11 * abs(lam) + 1e-300
    assert np.abs(back["vec"] - v).max() <= 1e-11


def test_fiedler_ibm10_residual(handles, oracle, circuits):
    """ibm10's golden file is not a converged pair: check by residual and vs the oracle."""
    h = handles["ibm10"]
    h.assemble_laplacian()
    lam, v = h.fiedler()
    assert abs(lam - 0.0185035852) / 0.0185035852 < 1e-7                  # converged lambda2, SURVEY Appendix D
    assert np.linalg.norm(h.spmv(v) - lam * v) < 1e-9


def test_fiedler_ibm10_vs_independent_solve(oracle, circuits):
    """ibm10's Fiedler VECTOR against an independent converged solve: the oracle port's plain restarted Lanczos on the
    CPU (different algorithm -- no polynomial filter --, different code, different arithmetic order)."""
    with api.Handle() as h:
        h.load_hgr(circuits["ibm10"])
        h.assemble_laplacian()
        lam, v = h.fiedler()
    e = oracle.OracleEIG(oracle.OracleHgr(circuits["ibm10"]))
    lo, vo, st = e.fiedler()
    assert st["converged"] == 1
    assert np.linalg.norm(e.spmv(vo) - lo * vo) < 1e-8                   # the independent pair is a genuine eigenpair
    assert abs(lam - lo) / lo <= 1e-8                                    # north star: 1e-8 relative
    cs = abs(v @ vo)
    assert np.sqrt(max(0.0, 1.0 - cs * cs)) <= 1e-6                      # north star: 1e-6 sine


def test_fiedler_sign_is_canonical(circuits):
    """The returned vector's largest-magnitude component is positive, whatever the start vector: the side labels
    (and the fused EIG -> KL trace) do not depend on the seed or on the number of ranks."""
    vs = []
    for seed in (0, 1, 12345):
        with api.Handle(seed=seed) as h:
            h.load_hgr(circuits["ibm01"])
            h.assemble_laplacian()
            lam, v = h.fiedler()
            vs.append(v)
        assert v[np.argmax(np.abs(v))] > 0
    assert v @ vs[0] > 0.999999 and vs[1] @ vs[0] > 0.999999


def test_fiedler_is_reproducible(handles):
    h = handles["ibm01"]
    h.assemble_laplacian()
    l1, v1 = h.fiedler()
    l2, v2 = h.fiedler()
    assert l1 == l2 and np.array_equal(v1, v2)                            # fixed-order reductions


def test_fused_pipeline_equals_file_handoff(handles, oracle, circuits, tmp_path):
    """EIG -> KL in one process (gKL2.cu:1018-1024's aim) == cEIG file -> cKL -EIG, and == the oracle."""
    c = "ibm01"
    h = handles[c]
    h.assemble_laplacian()
    h.fiedler(want_vector=False)
    med, side = h.partition_from_fiedler()
    h.assemble_kl_graph()
    tr = h.kl_run()
    o = oracle.OracleKL(oracle.OracleHgr(circuits[c]))
    ro = o.run(side)
    assert tr["swaps"] == ro["swaps"]
    assert np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["node2"], ro["node2"])
    assert np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))
    out = str(tmp_path / "e.txt")
    h.write_eig(out)
    h.load_eig(out)
    tr2 = h.kl_run()
    assert np.array_equal(tr2["node1"], tr["node1"]) and np.array_equal(tr2["cut"], tr["cut"])


def test_plain_lanczos_flag_matches_filtered(handles, oracle, workdir):
    """EIGKL_F_PLAIN_LANCZOS (Lanczos on L itself, the Spectra-style iteration) and the default Chebyshev-filtered
    solver must agree with each other and with the golden file."""
    c = "ibm01"
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), handles[c].n_nodes)
    out = {}
    for name, flags in (("filtered", 0), ("plain", api.EIGKL_F_PLAIN_LANCZOS)):
        with api.Handle(flags=flags) as h:
            h.load_hgr(os.path.join(workdir, "circuit", c + ".hgr"))
            h.assemble_laplacian()
            lam, v = h.fiedler()
            st = h.stats()
            out[name] = (lam, v, st)
            assert st["converged"] == 1 and st["cheb_degree"] == (1 if flags else 16)
            assert abs(lam - g["lambda2"]) / g["lambda2"] <= 1e-8
            cs = abs(v @ g["vec"]) / np.linalg.norm(g["vec"])
            assert np.sqrt(max(0.0, 1.0 - cs * cs)) <= 1e-6
    assert abs(out["plain"][0] - out["filtered"][0]) <= 1e-10 * out["plain"][0]
    assert abs(abs(out["plain"][1] @ out["filtered"][1]) - 1.0) < 1e-12
    assert out["filtered"][2]["lanczos_steps"] * 4 < out["plain"][2]["lanczos_steps"]     # the point of the filter


@pytest.mark.parametrize("c,env_k", [("fract", None), ("ibm01", None), ("ibm01", "16"), ("ibm01", "24"), ("industry2", None),
                                     ("industry2", "16"), ("ibm10", None)])
def test_resident_filter_matches_per_spmv_launches(c, env_k, workdir, monkeypatch):
    """The resident polynomial filter (a whole Chebyshev filter application as one cooperative launch: matrix in
    registers, x in shared memory, halo exchanged through L2) against the path with one kernel launch per SpMV,
    in every entries-per-thread variant of the kernel."""
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("EIGKL_SPMV_RESIDENT", mode)
        if env_k:
            monkeypatch.setenv("EIGKL_RES_K", env_k)
        with api.Handle() as h:
            h.load_hgr(os.path.join(workdir, "circuit", c + ".hgr"))
            h.assemble_laplacian()
            lam, v = h.fiedler()
            lam2, v2 = h.fiedler()
            assert lam == lam2 and np.array_equal(v, v2)                  # no order dependence in the halo exchange
            r = np.linalg.norm(h.spmv(v) - lam * v)
            res[mode] = (lam, v, h.stats(), r)
    s0, s1 = res["0"][2], res["1"][2]
    assert s0["resident_k"] == 0 and s0["spmv_per_launch"] == 1
    assert s1["spmv_per_launch"] == 16 and s1["resident_k"] == (int(env_k) if env_k else s1["resident_k"])
    assert s1["resident_k"] in (4, 8, 16, 24)
    assert s0["converged"] == 1 and s1["converged"] == 1
    assert s1["matvecs"] == s0["matvecs"]                                 # same iteration, different plumbing
    assert s1["gpu_launches"] < s0["gpu_launches"]
    assert abs(res["0"][0] - res["1"][0]) <= 1e-10 * res["0"][0]
    cs = abs(res["0"][1] @ res["1"][1])
    assert np.sqrt(max(0.0, 1.0 - cs * cs)) <= 1e-6
    assert res["0"][3] < 1e-9 and res["1"][3] < 1e-9


@pytest.mark.parametrize("c,cache", [("fract", None), ("ibm01", None), ("ibm01", "7"), ("ibm01", "0"), ("ibm10", None)])
def test_fused_gram_schmidt_matches_separate_kernels(c, cache, workdir, monkeypatch):
    """Both Gram-Schmidt passes of a Lanczos step as one cooperative launch (basis slice cached in shared memory,
    grid barriers between the reductions) against the separate multidot / update launches; `cache` limits the
    number of basis columns kept on chip so that the re-read-from-L2 path runs too."""
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("EIGKL_GS_FUSED", mode)
        if cache is not None:
            monkeypatch.setenv("EIGKL_GS_CACHE", cache)
        with api.Handle() as h:
            h.load_hgr(os.path.join(workdir, "circuit", c + ".hgr"))
            h.assemble_laplacian()
            lam, v = h.fiedler()
            lam2, v2 = h.fiedler()
            assert lam == lam2 and np.array_equal(v, v2)                  # fixed-order folds: bit-reproducible
            res[mode] = (lam, v, h.stats(), np.linalg.norm(h.spmv(v) - lam * v))
    s0, s1 = res["0"][2], res["1"][2]
    assert s0["gs_fused"] == 0 and s1["gs_fused"] == 1
    if cache is not None:
        assert s1["gs_cache_cols"] == int(cache)
    assert s0["converged"] == 1 and s1["converged"] == 1
    assert s1["lanczos_steps"] == s0["lanczos_steps"] and s1["gpu_launches"] < s0["gpu_launches"]
    assert abs(res["0"][0] - res["1"][0]) <= 1e-10 * res["0"][0]
    cs = abs(res["0"][1] @ res["1"][1])
    assert np.sqrt(max(0.0, 1.0 - cs * cs)) <= 1e-6
    assert res["0"][3] < 1e-9 and res["1"][3] < 1e-9


@pytest.mark.parametrize("n", [8, 12, 31])
def test_tiny_graphs(n, oracle, tmp_path):
    """Smallest sizes the reference's ncv = min(100, n/2) rule allows: a ring of 2-pin nets plus one chord net."""
    nets = [[i, (i + 1) % n] for i in range(n)] + [[0, n // 2, n // 3]]
    path = str(tmp_path / "ring.hgr")
    with open(path, "w") as f:
        f.write(f"{len(nets)} {n}\n")
        for e in nets:
            f.write(" ".join(str(p + 1) for p in e) + "\n")
    o = oracle.OracleEIG(oracle.OracleHgr(path))
    L = np.zeros((n, n))
    for r in range(n):
        for e in range(o.rowptr[r], o.rowptr[r + 1]):
            L[r, o.col[e]] = o.val[e]
    w, V = np.linalg.eigh(L)
    with api.Handle() as h:
        h.load_hgr(path)
        h.assemble_laplacian()
        lam, v = h.fiedler()
        assert abs(lam - w[1]) <= 1e-9 * w[1]
        assert np.linalg.norm(L @ v - lam * v) < 1e-9
        med, side = h.partition_from_fiedler()
        h.assemble_kl_graph()
        tr = h.kl_run()
        ro = oracle.OracleKL(oracle.OracleHgr(path)).run(side)
        assert np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))
    with api.Handle() as h:                           # below the reference's minimum: refused, not mis-solved
        h.set_pins(4, np.array([0, 2, 4], np.int64), np.array([0, 1, 2, 3], np.int32))
        h.assemble_laplacian()
        with pytest.raises(api.EigklError) as e:
            h.fiedler()
        assert e.value.code == -1
