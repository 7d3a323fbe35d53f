"""Pins the CPU oracle (oracle/eigkl_oracle.c) to the reference.

KL: byte-exact against trace files produced by running the reference itself on one core
(tests/golden/make_golden.sh -> oracle/_ref/cKL, built from /root/reference/cKL.cpp), swap node ids
against its instrumented twin, md5 sums against SURVEY.md Appendix D.
EIG: against the reference's shipped golden outputs pre_saved_EIG/<c>.hgr_out.txt.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from eig_kl_algorithm_b200 import datasets

APPENDIX_D = {   # circuit: (swaps, full-md5 of the 1-core trace, gaincol-md5, swapseq-md5)
    "fract": (18, "9f9b518d58b38ae48c59e84e375b2a74", "c05a4c8061212486ed1c963d1a6e0477", "0343eb4d58eaa938154539bf10d0cfcb"),
    "ibm01": (164, "bc16f641b1cc972b029fba3999ccacd0", "fa38ff06bd3a32f8082d324fa26864a4", "11d6985797e97a84c008c135ff7bd23e"),
    "industry2": (115, "b0357fbfeaf61b7f084fb46379a747ff", "8584ff2b8aa2a34e1d747ffe232c3f66", "542862abd6dce8e4cd7254268117b439"),
    "ibm10": (1544, "5f8223e7cf82c941fcb2d9a726c53cb5", "0186edc0cda4492e683ab1d3e417b999", "afce4c4cbc83df6fea254e65e09dd421"),
}


def golden_swaps(c):
    a = np.loadtxt(os.path.join(GOLDEN, c + ".kl_swaps.txt"), skiprows=1, usecols=(3, 4), dtype=np.int64, ndmin=2)
    return a[:, 0], a[:, 1]


@pytest.mark.parametrize("c", list(APPENDIX_D))
def test_golden_files_match_survey(c):
    swaps, full, gaincol, swapseq = APPENDIX_D[c]
    raw = open(os.path.join(GOLDEN, c + ".kl_trace_1core.txt"), "rb").read()
    assert hashlib.md5(raw).hexdigest() == full
    rows = [r.split("\t") for r in raw.decode().splitlines()]
    assert len(rows) == swaps + 1
    assert hashlib.md5("".join(f"{r[0]}\t{r[2]}\n" for r in rows).encode()).hexdigest() == gaincol
    n1, n2 = golden_swaps(c)
    assert hashlib.md5("".join(f"{a}\t{b}\n" for a, b in zip(n1, n2)).encode()).hexdigest() == swapseq


@pytest.mark.parametrize("c", list(APPENDIX_D))
def test_oracle_kl_trace_byte_exact(c, oracle, workdir, circuits, tmp_path):
    h = oracle.OracleHgr(circuits[c])
    kl = oracle.OracleKL(h)
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), h.n_nodes)
    r = kl.run(g["side"])
    out = str(tmp_path / "trace.txt")
    oracle.write_trace(out, r["cut"], r["gain"])
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, c + ".kl_trace_1core.txt"), "rb").read()
    n1, n2 = golden_swaps(c)
    assert r["swaps"] == APPENDIX_D[c][0]
    assert np.array_equal(r["node1"][1:], n1) and np.array_equal(r["node2"][1:], n2)
    # the final partition is balanced exactly like the initial one (pairwise swaps)
    assert int(r["side"].sum()) == int(g["side"].sum())


@pytest.mark.parametrize("c", ["fract", "ibm01", "industry2"])
def test_oracle_block_cached_selection_equals_linear_scan(c, oracle, workdir, circuits):
    # orc_kl_run caches the first best element per 256 positions of remain[]; the literal scans of
    # cKL.cpp:341-355 (orc_kl_run_linear) must give the same pass, from the golden start and from shuffled orders
    h = oracle.OracleHgr(circuits[c])
    kl = oracle.OracleKL(h)
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), h.n_nodes)
    rng = np.random.default_rng(7)
    perm = rng.permutation(h.n_nodes).astype(np.int32)
    half = h.n_nodes // 2
    starts = [(g["side"], None, None)]
    side = np.zeros(h.n_nodes, np.uint8)
    side[perm[half:]] = 1
    starts.append((side, perm[:half], perm[half:]))
    for sd, o0, o1 in starts:
        ra, rb = kl.run(sd, o0, o1), kl.run(sd, o0, o1, linear=True)
        assert ra["swaps"] == rb["swaps"]
        for k in ("node1", "node2", "side"):
            assert np.array_equal(ra[k], rb[k])
        for k in ("cut", "gain"):
            assert np.array_equal(ra[k].view(np.uint32), rb[k].view(np.uint32))


def test_oracle_kl_label_swap_invariance(oracle, workdir, circuits):
    # flipping every side bit leaves the gain column and the swap pairs unchanged
    c = "ibm01"
    h = oracle.OracleHgr(circuits[c])
    kl = oracle.OracleKL(h)
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), h.n_nodes)
    r0 = kl.run(g["side"])
    r1 = kl.run(1 - g["side"])
    assert np.array_equal(r0["gain"], r1["gain"])
    assert np.array_equal(r0["node1"][1:], r1["node2"][1:]) and np.array_equal(r0["node2"][1:], r1["node1"][1:])


def _fiedler_checks(oracle, workdir, circuits, c):
    h = oracle.OracleHgr(circuits[c])
    e = oracle.OracleEIG(h)
    g = oracle.read_eig(datasets.golden_eig_path(workdir, c), h.n_nodes)
    lam, v, st = e.fiedler()
    return h, e, g, lam, v, st


@pytest.mark.parametrize("c", ["fract", "ibm01"])
def test_oracle_eig_vs_golden(c, oracle, workdir, circuits, tmp_path):
    h, e, g, lam, v, st = _fiedler_checks(oracle, workdir, circuits, c)
    assert st["converged"] == 1
    assert abs(lam - g["lambda2"]) / g["lambda2"] <= 1e-8            # north-star tolerance
    cs = abs(v @ g["vec"]) / np.linalg.norm(g["vec"])
    assert np.sqrt(max(0.0, 1 - cs * cs)) <= 1e-6                     # sine after sign alignment
    s = np.sign(v @ g["vec"])
    # side column = (median > v_i), cEIG.cpp:218
    side = (oracle.median(v * s) > v * s).astype(np.uint8)
    assert int((side != g["side"]).sum()) == 0
    # writer: same format as the golden file (header values to print precision, identical side column)
    out = str(tmp_path / "eig.txt")
    oracle.write_eig(out, lam, v * s)
    got = open(out).read().splitlines()
    want = open(datasets.golden_eig_path(workdir, c)).read().splitlines()
    assert len(got) == len(want)
    assert [r.split("\t")[:2] for r in got[2:]] == [r.split("\t")[:2] for r in want[2:]]
    assert abs(float(got[1]) - float(want[1])) <= 1e-9


def test_oracle_laplacian_properties(oracle, circuits):
    h = oracle.OracleHgr(circuits["ibm01"])
    e = oracle.OracleEIG(h)
    n = e.n
    assert int(e.rowptr[-1]) == 231118                                 # nnz(L) of ibm01
    y = e.spmv(np.ones(n))
    assert np.abs(y).max() < 1e-12                                     # L 1 = 0
    # symmetric: x^T L y == y^T L x
    rng = np.random.default_rng(0)
    x, z = rng.standard_normal(n), rng.standard_normal(n)
    assert abs(x @ e.spmv(z) - z @ e.spmv(x)) < 1e-9


def test_oracle_sym_eig(oracle):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((80, 80))
    a = a + a.T
    d, z = oracle.sym_eig(a)
    assert np.allclose(d, np.linalg.eigvalsh(a), atol=1e-11)
    assert np.abs(a @ z - z * d).max() < 1e-11
