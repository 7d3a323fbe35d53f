"""Full-size checks (BASELINE.json configs 4-5: ibm18-sized and 2 M-node synthetic circuits) through
size-independent properties, plus bit-exact parity with the oracle at the ibm18-sized circuit.

The synthetic circuits are disconnected: lambda2 = 0, so the EIG checks are residual based.
"""
import os

import numpy as np
import pytest

from eig_kl_algorithm_b200 import api, datasets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth1(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    return datasets.write_synthetic(str(d / "synth1.hgr"), 1.0, seed=12345)


def _kl_invariants(h, tr, side0):
    # incremental cut bookkeeping is exact fp32: cut[i] = cut[i-1] - gain[i]        (cKL.cpp:362)
    assert np.array_equal(tr["cut"][1:], (tr["cut"][:-1] - tr["gain"][1:]).astype(np.float32))
    # each node is swapped at most once, node1 leaves side 0 and node2 leaves side 1   (cKL.cpp:274-286)
    n1, n2 = tr["node1"][1:], tr["node2"][1:]
    assert len(np.unique(n1)) == len(n1) and len(np.unique(n2)) == len(n2)
    assert np.all(side0[n1] == 0) and np.all(side0[n2] == 1)
    side = h.get_partition()
    expect = side0.copy()
    expect[n1] = 1
    expect[n2] = 0
    assert np.array_equal(side, expect)
    assert int(side.sum()) == int(side0.sum())                 # balance is preserved by pairwise swaps
    # the D-values the loop maintained incrementally == a from-scratch pass over the final partition
    kept = h.kl_values()
    h.set_partition(side)
    assert np.array_equal(kept.view(np.uint32), h.dvalues().view(np.uint32))
    # the final cut recomputed from scratch agrees with the running value up to fp32 drift: both are the
    # reference's single-accumulator float sums (cKL.cpp:199-223, 362) over 1e5..1e6 terms at magnitude ~5e5
    # (ulp 0.03-0.06), so only a few per cent can be asked; the exact check is the D-value equality above
    assert abs(float(h.cut()) - float(tr["cut"][-1])) <= 3e-2 * max(1.0, abs(float(tr["cut"][-1])))
    # termination rule: the pass ends after floor(log2 N)+6 consecutive non-positive gains (cKL.cpp:303,382-386)
    limit = int(np.log2(h.n_nodes)) + 5
    tail = tr["gain"][-(limit + 1):]
    if tr["swaps"] < min((side0 == 0).sum(), (side0 == 1).sum()):
        assert np.all(tail <= 0) and tr["gain"][-(limit + 2)] > 0


def test_ibm18_sized_synthetic_bit_exact_vs_oracle(synth1, oracle):
    """201 920 nodes / 210 613 nets (pins 522 157, pairs 550 934): GPU == oracle, bit for bit."""
    n, off, pins = datasets.read_hgr_arrays(synth1)
    k = np.diff(off)
    assert (n, len(off) - 1, len(pins), int((k * (k - 1) // 2).sum())) == (201920, 210613, 522157, 550934)
    side0 = np.random.default_rng(3).integers(0, 2, n).astype(np.uint8)
    with api.Handle() as h:
        h.set_pins(n, off, pins)
        h.assemble_kl_graph()
        rp, fe, col, w = h.get_kl_graph()
        o = oracle.OracleKL(oracle.OracleHgr(synth1))
        assert np.array_equal(rp, o.rowptr.astype(np.int32)) and np.array_equal(col, o.col)
        assert np.array_equal(w.view(np.uint32), o.w.view(np.uint32))
        h.set_partition(side0)
        assert np.array_equal(h.dvalues().view(np.uint32), o.dvalues(side0).view(np.uint32))
        assert h.cut().view(np.uint32) == o.cut0(side0).view(np.uint32)
        tr = h.kl_run()
        ro = o.run(side0)
        assert tr["swaps"] == ro["swaps"]
        assert np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["node2"], ro["node2"])
        assert np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))
        h.set_partition(side0)
        tr = h.kl_run()
        _kl_invariants(h, tr, side0)


def test_ibm18_sized_synthetic_eig_properties(synth1):
    with api.Handle() as h:
        h.load_hgr(synth1)
        h.assemble_laplacian()
        n = h.n_nodes
        rng = np.random.default_rng(0)
        x, z = rng.standard_normal(n), rng.standard_normal(n)
        lx, lz = h.spmv(x), h.spmv(z)
        assert np.abs(h.spmv(np.ones(n))).max() < 1e-10            # L 1 = 0
        assert abs(x @ lz - z @ lx) < 1e-8 * abs(x @ lz)           # symmetry
        assert x @ lx > 0                                          # positive semi-definite
        assert np.allclose(h.spmv(2.0 * x - 3.0 * z), 2.0 * lx - 3.0 * lz, rtol=1e-12, atol=1e-9)   # linearity
        lam, v = h.fiedler()
        st = h.stats()
        assert abs(np.linalg.norm(v) - 1) < 1e-12
        assert abs(lam) < 1e-10                                    # disconnected: second eigenvalue is 0 too
        assert np.linalg.norm(h.spmv(v) - lam * v) < 1e-9          # a genuine eigenpair
        assert st["resid_est"][1] < 1e-9
        # 201 920 nodes / 1.2 M entries still fit on chip: the resident filter and the fused Gram-Schmidt kernel run here
        assert st["spmv_per_launch"] == (16 if st["resident_k"] else 1) and st["gs_fused"] == 1
        med, side = h.partition_from_fiedler()
        assert np.array_equal(side, (med > v).astype(np.uint8))    # cEIG.cpp:218
        s = np.sort(v)
        assert med == (s[(n - 1) // 2] + s[n // 2]) / 2.0 if n % 2 == 0 else s[n // 2]   # cEIG.cpp:55-65


@pytest.mark.skipif(os.environ.get("EIGKL_SKIP_2M") == "1", reason="2M-node case skipped by request")
def test_two_million_node_synthetic_properties(tmp_path, oracle):
    """BASELINE.json config 5 (circuit_generator scale 10, ~2 M nodes): the whole fused pipeline, checked by
    invariants AND, for the KL pass, bit for bit against the oracle (its block-cached selection makes the 280 K-swap
    pass a ~10 s job; tests/test_oracle.py pins that selection to the literal scans and to the reference's traces)."""
    path = datasets.write_synthetic(str(tmp_path / "synth10.hgr"), 10.0, seed=12345)
    with api.Handle() as h:
        h.load_hgr(path)
        assert h.n_nodes == 2019200 and h.n_nets == 2106130
        h.assemble_laplacian()
        lam, v = h.fiedler()
        assert abs(lam) < 1e-9 and np.linalg.norm(h.spmv(v) - lam * v) < 1e-8
        st = h.stats()                                                       # too large to stay on chip:
        assert st["resident_k"] == 0 and st["spmv_per_launch"] == 1 and st["gs_fused"] == 0   # the streaming kernels run
        med, side0 = h.partition_from_fiedler()
        assert abs(int(side0.sum()) - h.n_nodes // 2) <= h.n_nodes // 2     # any split is legal for a null vector
        h.assemble_kl_graph()
        tr = h.kl_run()
        assert h.stats()["kl_local"] == 2                       # one CTA: tile keys in shared memory, state bytes in global memory
        assert tr["swaps"] > 1000
        assert tr["cut"].min() <= tr["cut"][0]
        o = oracle.OracleKL(oracle.OracleHgr(path))
        ro = o.run(side0)
        assert tr["swaps"] == ro["swaps"]
        assert np.array_equal(tr["node1"], ro["node1"]) and np.array_equal(tr["node2"], ro["node2"])
        assert np.array_equal(tr["cut"].view(np.uint32), ro["cut"].view(np.uint32))
        assert np.array_equal(tr["gain"].view(np.uint32), ro["gain"].view(np.uint32))
        _kl_invariants(h, tr, side0)
