"""Host-side logic of the product on the CPU: text formats, dense eigen-solver, the C ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from eig_kl_algorithm_b200 import api, datasets
from helpers import build_helpers


@pytest.fixture(scope="module")
def shim():
    L = C.CDLL(build_helpers.build())
    P = C.POINTER
    L.shim_parse_hgr.argtypes = [C.c_char_p, P(C.c_int32), P(C.c_int32), P(C.c_int64), C.c_int64, P(C.c_int32), C.c_int64, C.c_char_p, C.c_int]
    L.shim_write_eig.argtypes = [C.c_char_p, C.c_double, C.c_double, P(C.c_double), C.c_int32]
    L.shim_read_eig.argtypes = [C.c_char_p, C.c_int32, P(C.c_uint8), C.c_char_p, C.c_int]
    L.shim_read_eig_orders.argtypes = [C.c_char_p, C.c_int32, P(C.c_uint8), P(C.c_int32), P(C.c_int32)]
    L.shim_sym_eig.argtypes = [C.c_int, P(C.c_double), P(C.c_double)]
    L.shim_tridiag_top.argtypes = [C.c_int, P(C.c_double), P(C.c_double), C.c_int, P(C.c_double), P(C.c_double)]
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _parse(shim, path, cap=4_000_000):
    nn, ne = C.c_int32(), C.c_int32()
    off = np.zeros(cap, np.int64)
    pins = np.zeros(cap, np.int32)
    err = C.create_string_buffer(256)
    rc = shim.shim_parse_hgr(path.encode(), C.byref(nn), C.byref(ne), _p(off, C.c_int64), cap, _p(pins, C.c_int32), cap, err, 256)
    return rc, nn.value, ne.value, off, pins, err.value.decode()


@pytest.mark.parametrize("c", ["fract", "ibm01", "industry2"])
def test_parser_matches_oracle(c, shim, oracle, circuits):
    rc, nn, ne, off, pins, _ = _parse(shim, circuits[c])
    assert rc == 0
    h = oracle.OracleHgr(circuits[c])
    assert (nn, ne) == (h.n_nodes, h.n_nets)
    assert np.array_equal(off[: ne + 1], h.net_off)
    assert np.array_equal(pins[: off[ne]], h.pins)


def test_parser_edge_cases(shim, tmp_path):
    p = tmp_path / "t.hgr"
    # header with a third token, trailing blanks, an empty net line, a 1-pin net, missing trailing lines
    p.write_text("5 6 1\n1 2 3 \n\n4\n 5   6\t2\n")
    rc, nn, ne, off, pins, _ = _parse(shim, str(p))
    assert rc == 0 and (nn, ne) == (6, 5)
    assert list(off[:6]) == [0, 3, 3, 4, 7, 7]
    assert list(pins[:7]) == [0, 1, 2, 3, 4, 5, 1]
    p.write_text("1 3\n1 4\n")                           # pin id > nodes
    assert _parse(shim, str(p))[0] == -3
    p.write_text("1 3\n0 2\n")                           # pin id 0
    assert _parse(shim, str(p))[0] == -3
    p.write_text("x y\n")
    assert _parse(shim, str(p))[0] == -3
    assert _parse(shim, str(tmp_path / "missing.hgr"))[0] == -2


@pytest.mark.parametrize("c", ["fract", "ibm01"])
def test_eig_writer_reproduces_golden_file(c, shim, oracle, workdir, tmp_path):
    # the golden file's own numbers pushed through our writer must give the golden file back, byte for byte
    path = datasets.golden_eig_path(workdir, c)
    lines = open(path).read().splitlines()
    n = len(lines) - 2
    g = oracle.read_eig(path, n)
    out = str(tmp_path / "o.txt")
    assert shim.shim_write_eig(out.encode(), g["lambda2"], g["median"], _p(g["vec"], C.c_double), n) == 0
    assert open(out, "rb").read() == open(path, "rb").read()
    side = np.zeros(n, np.uint8)
    err = C.create_string_buffer(256)
    assert shim.shim_read_eig(out.encode(), n, _p(side, C.c_uint8), err, 256) == 0
    assert np.array_equal(side, g["side"])


def test_parallel_text_io_is_thread_count_independent(shim, oracle, circuits, workdir, tmp_path, monkeypatch):
    # the parser / EIG writer / EIG reader split their input at newlines, one piece per thread: any thread count
    # must give the same arrays and the same bytes (EIGKL_IO_THREADS overrides the automatic choice)
    c = "ibm01"
    h = oracle.OracleHgr(circuits[c])
    gpath = datasets.golden_eig_path(workdir, c)
    g = oracle.read_eig(gpath, h.n_nodes)
    for threads in ("1", "2", "7", "16"):
        monkeypatch.setenv("EIGKL_IO_THREADS", threads)
        rc, nn, ne, off, pins, _ = _parse(shim, circuits[c])
        assert rc == 0 and (nn, ne) == (h.n_nodes, h.n_nets)
        assert np.array_equal(off[: ne + 1], h.net_off) and np.array_equal(pins[: off[ne]], h.pins)
        out = str(tmp_path / f"o{threads}.txt")
        assert shim.shim_write_eig(out.encode(), g["lambda2"], g["median"], _p(g["vec"], C.c_double), h.n_nodes) == 0
        assert open(out, "rb").read() == open(gpath, "rb").read()
        side = np.zeros(h.n_nodes, np.uint8)
        order = np.zeros(h.n_nodes, np.int32)
        n0 = C.c_int32()
        assert shim.shim_read_eig_orders(out.encode(), h.n_nodes, _p(side, C.c_uint8), _p(order, C.c_int32), C.byref(n0)) == 1
        assert np.array_equal(side, g["side"]) and n0.value == int((g["side"] == 0).sum())
        assert np.array_equal(order[: n0.value], np.nonzero(g["side"] == 0)[0])
    # a ragged file: no trailing newline, CRLF, blank lines, lines beyond <nets>, a non-numeric token ending a net
    p = tmp_path / "r.hgr"
    p.write_text("4 9\r\n1 2 3\r\n\r\n4 x 5\n6 7\n8 9\n1 2")
    for threads in ("1", "3", "5"):
        monkeypatch.setenv("EIGKL_IO_THREADS", threads)
        rc, nn, ne, off, pins, _ = _parse(shim, str(p))
        assert rc == 0 and (nn, ne) == (9, 4)
        assert list(off[:5]) == [0, 3, 3, 4, 6] and list(pins[:6]) == [0, 1, 2, 3, 5, 6]


def test_eig_reader_keeps_file_order_of_unsorted_files(shim, tmp_path, monkeypatch):
    # cKL.cpp:166-173 appends nodes to remain[side] in FILE order; a file that is not ascending keeps that order
    p = tmp_path / "u.txt"
    p.write_text("0.1\n0.0\n3\t1\t0.5\n0\t0\t-0.5\n2\t1\t0\n1\t0\t0.25\n4\t0\t0.1\n")
    for threads in ("1", "3"):
        monkeypatch.setenv("EIGKL_IO_THREADS", threads)
        side = np.zeros(5, np.uint8)
        order = np.zeros(5, np.int32)
        n0 = C.c_int32()
        assert shim.shim_read_eig_orders(str(p).encode(), 5, _p(side, C.c_uint8), _p(order, C.c_int32), C.byref(n0)) == 0
        assert list(side) == [0, 0, 1, 1, 0] and n0.value == 3 and list(order) == [0, 1, 4, 3, 2]
    p.write_text("0.1\n0.0\n0\t1\t0.5\n0\t0\t-0.5\n1\t1\t0\n")       # node 0 twice
    assert shim.shim_read_eig_orders(str(p).encode(), 2, _p(side, C.c_uint8), _p(order, C.c_int32), C.byref(n0)) == -3


def test_eig_reader_errors(shim, tmp_path):
    err = C.create_string_buffer(256)
    side = np.zeros(3, np.uint8)
    assert shim.shim_read_eig(str(tmp_path / "nope.txt").encode(), 3, _p(side, C.c_uint8), err, 256) == -2
    assert b"EIG file not found" in err.value                         # cKL.cpp:158
    p = tmp_path / "e.txt"
    p.write_text("0.1\n0.0\n0\t1\t0.5\n1\t0\t-0.5\n")                 # 2 rows for 3 nodes
    assert shim.shim_read_eig(str(p).encode(), 3, _p(side, C.c_uint8), err, 256) == -3
    p.write_text("0.1\n0.0\n0\t1\t0.5\n1\t2\t-0.5\n2\t0\t0\n")        # side 2
    assert shim.shim_read_eig(str(p).encode(), 3, _p(side, C.c_uint8), err, 256) == -3


@pytest.mark.parametrize("n", [2, 3, 17, 74, 100])
def test_dense_eig(shim, n):
    rng = np.random.default_rng(n)
    # arrowhead + tridiagonal, the shape a thick restart produces
    k = min(5, n - 1)
    a = np.zeros((n, n))
    a[np.arange(n), np.arange(n)] = rng.standard_normal(n)
    a[k, :k] = a[:k, k] = rng.standard_normal(k) * 1e-3
    for j in range(k, n - 1):
        a[j, j + 1] = a[j + 1, j] = abs(rng.standard_normal())
    z = np.array(a, order="C", copy=True)
    d = np.zeros(n)
    assert shim.shim_sym_eig(n, _p(z, C.c_double), _p(d, C.c_double)) == 0
    assert np.allclose(d, np.linalg.eigvalsh(a), atol=1e-12 * max(1.0, np.abs(a).max()) * n)
    assert np.abs(a @ z - z * d).max() < 1e-12 * n
    assert np.abs(z.T @ z - np.eye(n)).max() < 1e-12 * n


@pytest.mark.parametrize("n,keep,shape", [(100, 20, "arrow"), (100, 20, "tridiag"), (37, 7, "arrow"), (30, 5, "dense"), (6, 2, "arrow")])
def test_sym_top_eig(shim, n, keep, shape):
    """The k largest pairs of the projected matrix of a thick-restart cycle (what the restart and the convergence
    check need) by Householder (skipping rows already tridiagonal) + bisection + inverse iteration, vs numpy."""
    rng = np.random.default_rng(7 * n + keep)
    a = np.zeros((n, n))
    if shape == "dense":
        b = rng.standard_normal((n, n))
        a = b + b.T
    else:
        kk = keep if shape == "arrow" else 0
        a[np.arange(n), np.arange(n)] = np.sort(np.abs(rng.standard_normal(n)) * np.logspace(3, 0, n))[::-1]
        if kk:
            a[kk, :kk] = a[:kk, kk] = rng.standard_normal(kk) * 1e-2
        for j in range(kk, n - 1):
            a[j, j + 1] = a[j + 1, j] = abs(rng.standard_normal()) + 0.05
    th = np.zeros(keep)
    Y = np.zeros(n * keep)
    shim.shim_sym_top_eig.restype = C.c_int
    assert shim.shim_sym_top_eig(C.c_int(n), _p(np.ascontiguousarray(a), C.c_double), C.c_int(keep), _p(th, C.c_double), _p(Y, C.c_double)) == 0
    w, V = np.linalg.eigh(a)
    scale = np.abs(w).max()
    assert np.allclose(th, w[::-1][:keep], atol=1e-12 * scale)
    Y = Y.reshape(keep, n).T
    assert np.abs(Y.T @ Y - np.eye(keep)).max() < 1e-10
    for t in range(keep):
        assert np.linalg.norm(a @ Y[:, t] - th[t] * Y[:, t]) < 1e-11 * scale


@pytest.mark.parametrize("n", [2, 3, 10, 37, 100])
def test_tridiag_top_eig(shim, n):
    """Top-k pairs of a Lanczos-like tridiagonal (bisection + inverse iteration) vs numpy."""
    rng = np.random.default_rng(100 + n)
    d = rng.standard_normal(n) * 3 + np.linspace(5, -5, n)
    e = np.abs(rng.standard_normal(max(n - 1, 1))) + 0.1
    T = np.diag(d) + np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1)
    k = min(3, n)
    th = np.zeros(k)
    Y = np.zeros(n * k)
    assert shim.shim_tridiag_top(n, _p(d, C.c_double), _p(e, C.c_double), k, _p(th, C.c_double), _p(Y, C.c_double)) == 0
    w, V = np.linalg.eigh(T)
    assert np.allclose(th, w[::-1][:k], atol=1e-12 * np.abs(w).max())
    Y = Y.reshape(k, n).T
    for t in range(k):
        assert np.linalg.norm(T @ Y[:, t] - th[t] * Y[:, t]) < 1e-11 * np.abs(w).max()
        assert abs(abs(Y[:, t] @ V[:, n - 1 - t]) - 1) < 1e-10
    assert np.abs(Y.T @ Y - np.eye(k)).max() < 1e-10


def test_partition_writer_format_and_thread_independence(shim, tmp_path, monkeypatch):
    """`--partition-out` file: one `node<TAB>side` row per node, ascending, whatever the writer's thread count."""
    rng = np.random.default_rng(5)
    n = 100003
    side = rng.integers(0, 2, n).astype(np.uint8)
    side[7] = 3                                   # only bit 0 is the side (bit 1 is the lock mark of a finished pass)
    shim.shim_write_partition.argtypes = [C.c_char_p, C.POINTER(C.c_uint8), C.c_int32]
    outs = []
    for threads in ("1", "7"):
        monkeypatch.setenv("EIGKL_IO_THREADS", threads)
        out = str(tmp_path / f"part_{threads}.txt")
        assert shim.shim_write_partition(out.encode(), _p(side, C.c_uint8), n) == 0
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]
    rows = outs[0].decode().split("\n")
    assert rows[-1] == "" and len(rows) == n + 1
    assert rows[0] == f"0\t{side[0] & 1}" and rows[7] == "7\t1" and rows[n - 1] == f"{n - 1}\t{side[n - 1] & 1}"
    assert shim.shim_write_partition(str(tmp_path / "no_such_dir" / "x.txt").encode(), _p(side, C.c_uint8), n) != 0


def test_trace_writer_format(eigkl_lib, tmp_path):
    # cKL.cpp:315,380 -- default ostream float formatting (6 significant digits)
    cut = np.array([27.75, 26.875, 36.958332, 1153.3374, 740.9452], np.float32)
    gain = np.array([0, 0.875, -0.958333, 1e-7, 1], np.float32)
    out = str(tmp_path / "t.txt")
    api.write_trace(out, dict(cut=cut, gain=gain))
    assert open(out).read() == "0\t27.75\t0\n1\t26.875\t0.875\n2\t36.9583\t-0.958333\n3\t1153.34\t1e-07\n4\t740.945\t1\n"


def test_cabi_exports_every_declared_symbol(eigkl_lib):
    header = open(os.path.join(ROOT, "include", "eigkl.h")).read()
    declared = sorted(set(re.findall(r"\b(eigkl_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(api.SYMBOLS)
    for s in declared:
        assert hasattr(eigkl_lib, s), s
    assert eigkl_lib.eigkl_abi_version() == 2


def test_ctypes_struct_layout_matches_header(tmp_path):
    # sizeof as the C compiler sees include/eigkl.h == sizeof of the ctypes mirrors
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "eigkl.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(eigkl_opts), sizeof(eigkl_stats), sizeof(eigkl_trace));return 0;}\n')
    exe = str(tmp_path / "sz")
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    o, s_, t = map(int, subprocess.check_output([exe]).split())
    assert (o, s_, t) == (C.sizeof(api.Opts), C.sizeof(api.Stats), C.sizeof(api.Trace))


def test_no_cpu_fallback(eigkl_lib):
    """Without a GPU every handle creation must fail loudly with E_CUDA (never a silent CPU path)."""
    if os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"):
        pytest.skip("GPU present")
    with pytest.raises(api.EigklError) as ei:
        api.Handle()
    assert ei.value.code == -4 and "no CPU fallback" in ei.value.message


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "eig_kl_algorithm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "eigkl_oracle" not in text, f
