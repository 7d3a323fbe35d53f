"""ctypes view of oracle/liboracle.so (the CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs.  Nothing under eig_kl_algorithm_b200/ may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")


class Hgr(C.Structure):
    _fields_ = [("n_nets", C.c_int32), ("n_nodes", C.c_int32),
                ("net_off", C.POINTER(C.c_int64)), ("pins", C.POINTER(C.c_int32))]


class KlGraph(C.Structure):
    _fields_ = [("n", C.c_int32), ("rowptr", C.POINTER(C.c_int64)), ("fwd_end", C.POINTER(C.c_int64)),
                ("col", C.POINTER(C.c_int32)), ("w", C.POINTER(C.c_float))]


class Csr(C.Structure):
    _fields_ = [("n", C.c_int32), ("rowptr", C.POINTER(C.c_int64)), ("col", C.POINTER(C.c_int32)),
                ("val", C.POINTER(C.c_double))]


class EigStats(C.Structure):
    _fields_ = [("matvecs", C.c_int32), ("restarts", C.c_int32), ("converged", C.c_int32),
                ("ncv", C.c_int32), ("resid_est", C.c_double * 2)]


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("eigkl_oracle.c", "eigkl_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.POINTER
        L.orc_hgr_load.argtypes = [C.c_char_p, P(Hgr)]
        L.orc_hgr_free.argtypes = [P(Hgr)]
        L.orc_stl_hash_order.argtypes = [P(C.c_uint32), C.c_int64, P(C.c_int64)]
        L.orc_kl_build.argtypes = [P(Hgr), P(KlGraph)]
        L.orc_kl_free.argtypes = [P(KlGraph)]
        L.orc_kl_dvalues.argtypes = [P(KlGraph), P(C.c_uint8), P(C.c_float)]
        L.orc_kl_cut0.argtypes = [P(KlGraph), P(C.c_uint8), P(C.c_int32), C.c_int64, P(C.c_int32), C.c_int64]
        L.orc_kl_cut0.restype = C.c_float
        L.orc_kl_run.argtypes = [P(KlGraph), P(C.c_uint8), P(C.c_int32), C.c_int64, P(C.c_int32), C.c_int64,
                                 P(C.c_float), P(C.c_float), P(C.c_int32), P(C.c_int32), C.c_int64]
        L.orc_kl_run.restype = C.c_int64
        L.orc_kl_run_linear.argtypes = L.orc_kl_run.argtypes
        L.orc_kl_run_linear.restype = C.c_int64
        L.orc_laplacian.argtypes = [P(Hgr), P(Csr)]
        L.orc_csr_free.argtypes = [P(Csr)]
        L.orc_spmv.argtypes = [P(Csr), P(C.c_double), P(C.c_double)]
        L.orc_fiedler.argtypes = [P(Csr), P(C.c_double), P(C.c_double), P(EigStats)]
        L.orc_fiedler_bounded.argtypes = [P(Csr), C.c_int, P(C.c_double), P(C.c_double), P(EigStats)]
        L.orc_median.argtypes = [P(C.c_double), C.c_int32]
        L.orc_median.restype = C.c_double
        L.orc_write_eig.argtypes = [C.c_char_p, C.c_double, P(C.c_double), C.c_int32]
        L.orc_read_eig_sides.argtypes = [C.c_char_p, C.c_int32, P(C.c_uint8), P(C.c_double), P(C.c_double), P(C.c_double)]
        L.orc_write_trace.argtypes = [C.c_char_p, P(C.c_float), P(C.c_float), C.c_int64]
        L.orc_sym_eig.argtypes = [C.c_int, P(C.c_double), P(C.c_double)]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleHgr:
    def __init__(self, path):
        self.h = Hgr()
        rc = lib().orc_hgr_load(path.encode(), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"orc_hgr_load({path}) -> {rc}")
        self.n_nets, self.n_nodes = self.h.n_nets, self.h.n_nodes
        self.net_off = np.ctypeslib.as_array(self.h.net_off, (self.n_nets + 1,)).copy()
        npins = int(self.net_off[-1])
        self.pins = np.ctypeslib.as_array(self.h.pins, (max(npins, 1),))[:npins].copy()

    def __del__(self):
        try:
            lib().orc_hgr_free(C.byref(self.h))
        except Exception:
            pass


class OracleKL:
    """KL graph + pass, reference semantics (cKL.cpp)."""

    def __init__(self, hgr: OracleHgr):
        self.hgr = hgr
        self.g = KlGraph()
        rc = lib().orc_kl_build(C.byref(hgr.h), C.byref(self.g))
        if rc != 0:
            raise RuntimeError(f"orc_kl_build -> {rc}")
        self.n = self.g.n
        self.rowptr = np.ctypeslib.as_array(self.g.rowptr, (self.n + 1,)).copy()
        nnz = int(self.rowptr[-1])
        self.fwd_end = np.ctypeslib.as_array(self.g.fwd_end, (max(self.n, 1),))[: self.n].copy()
        self.col = np.ctypeslib.as_array(self.g.col, (max(nnz, 1),))[:nnz].copy()
        self.w = np.ctypeslib.as_array(self.g.w, (max(nnz, 1),))[:nnz].copy()

    def __del__(self):
        try:
            lib().orc_kl_free(C.byref(self.g))
        except Exception:
            pass

    def dvalues(self, side):
        side = np.ascontiguousarray(side, dtype=np.uint8)
        out = np.empty(self.n, dtype=np.float32)
        lib().orc_kl_dvalues(C.byref(self.g), _p(side, C.c_uint8), _p(out, C.c_float))
        return out

    @staticmethod
    def _orders(side, order0, order1):
        if order0 is None:
            order0 = np.nonzero(side == 0)[0]
        if order1 is None:
            order1 = np.nonzero(side == 1)[0]
        return (np.ascontiguousarray(order0, dtype=np.int32), np.ascontiguousarray(order1, dtype=np.int32))

    def cut0(self, side, order0=None, order1=None):
        side = np.ascontiguousarray(side, dtype=np.uint8)
        o0, o1 = self._orders(side, order0, order1)
        return np.float32(lib().orc_kl_cut0(C.byref(self.g), _p(side, C.c_uint8), _p(o0, C.c_int32), len(o0),
                                            _p(o1, C.c_int32), len(o1)))

    def run(self, side, order0=None, order1=None, linear=False):
        """One KL pass.  linear=True: the literal O(|remain|) selection scans of cKL.cpp:341-355 instead of the
        block-cached selection (identical results; only feasible up to ~100 K nodes)."""
        side = np.array(side, dtype=np.uint8, copy=True)
        o0, o1 = self._orders(side, order0, order1)
        cap = min(len(o0), len(o1)) + 1
        cut = np.zeros(cap, np.float32)
        gain = np.zeros(cap, np.float32)
        n1 = np.zeros(cap, np.int32)
        n2 = np.zeros(cap, np.int32)
        fn = lib().orc_kl_run_linear if linear else lib().orc_kl_run
        swaps = fn(C.byref(self.g), _p(side, C.c_uint8), _p(o0, C.c_int32), len(o0),
                                 _p(o1, C.c_int32), len(o1), _p(cut, C.c_float), _p(gain, C.c_float),
                                 _p(n1, C.c_int32), _p(n2, C.c_int32), cap)
        s = int(swaps) + 1
        return dict(swaps=int(swaps), cut=cut[:s], gain=gain[:s], node1=n1[:s], node2=n2[:s], side=side)


class OracleEIG:
    def __init__(self, hgr: OracleHgr):
        self.hgr = hgr
        self.L = Csr()
        rc = lib().orc_laplacian(C.byref(hgr.h), C.byref(self.L))
        if rc != 0:
            raise RuntimeError(f"orc_laplacian -> {rc}")
        self.n = self.L.n
        self.rowptr = np.ctypeslib.as_array(self.L.rowptr, (self.n + 1,)).copy()
        nnz = int(self.rowptr[-1])
        self.col = np.ctypeslib.as_array(self.L.col, (nnz,)).copy()
        self.val = np.ctypeslib.as_array(self.L.val, (nnz,)).copy()

    def __del__(self):
        try:
            lib().orc_csr_free(C.byref(self.L))
        except Exception:
            pass

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n, np.float64)
        lib().orc_spmv(C.byref(self.L), _p(x, C.c_double), _p(y, C.c_double))
        return y

    def fiedler(self, max_restarts=0):
        lam = C.c_double()
        vec = np.empty(self.n, np.float64)
        st = EigStats()
        rc = lib().orc_fiedler_bounded(C.byref(self.L), max_restarts, C.byref(lam), _p(vec, C.c_double), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"orc_fiedler -> {rc}")
        return lam.value, vec, dict(matvecs=st.matvecs, restarts=st.restarts, converged=st.converged, ncv=st.ncv,
                                    resid_est=(st.resid_est[0], st.resid_est[1]))


def stl_hash_order(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.empty(len(keys), np.int64)
    lib().orc_stl_hash_order(_p(keys, C.c_uint32), len(keys), _p(out, C.c_int64))
    return out


def median(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_median(_p(v, C.c_double), len(v))


def write_eig(path, lam, vec):
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    rc = lib().orc_write_eig(path.encode(), lam, _p(vec, C.c_double), len(vec))
    if rc != 0:
        raise RuntimeError(f"orc_write_eig -> {rc}")


def read_eig(path, n):
    side = np.zeros(n, np.uint8)
    vec = np.zeros(n, np.float64)
    lam = C.c_double()
    med = C.c_double()
    rc = lib().orc_read_eig_sides(path.encode(), n, _p(side, C.c_uint8), C.byref(lam), C.byref(med), _p(vec, C.c_double))
    if rc != 0:
        raise RuntimeError(f"orc_read_eig_sides({path}) -> {rc}")
    return dict(lambda2=lam.value, median=med.value, side=side, vec=vec)


def write_trace(path, cut, gain):
    cut = np.ascontiguousarray(cut, dtype=np.float32)
    gain = np.ascontiguousarray(gain, dtype=np.float32)
    rc = lib().orc_write_trace(path.encode(), _p(cut, C.c_float), _p(gain, C.c_float), len(cut) - 1)
    if rc != 0:
        raise RuntimeError(f"orc_write_trace -> {rc}")


def sym_eig(a):
    a = np.array(a, dtype=np.float64, order="C", copy=True)
    n = a.shape[0]
    d = np.empty(n, np.float64)
    lib().orc_sym_eig(n, _p(a, C.c_double), _p(d, C.c_double))
    return d, a
