"""ctypes binding of libeigkl.so (include/eigkl.h) plus thin Python mirrors of the reference's
three executables (cEIG / cKL / gKL) for tests and bench.py.

There is no CPU fallback: if the shared library is missing this module raises at import of the
library, and every compute entry point fails loudly without a B200.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libeigkl.so")
BIN_DIR = os.path.join(PKG, "bin")

EIGKL_F_PROFILE = 0x1
EIGKL_F_PLAIN_LANCZOS = 0x4
EIGKL_F_NATURAL_ORDER = 0x8

ERRORS = {0: "OK", -1: "E_ARG", -2: "E_IO", -3: "E_FORMAT", -4: "E_CUDA", -5: "E_NCCL", -6: "E_NOCONV", -7: "E_NOMEM"}

# every symbol include/eigkl.h declares (checked by tests/test_cabi.py against the header itself)
SYMBOLS = [
    "eigkl_abi_version", "eigkl_nccl_unique_id", "eigkl_create", "eigkl_destroy", "eigkl_last_error",
    "eigkl_get_stats", "eigkl_synchronize", "eigkl_set_profile", "eigkl_load_hgr", "eigkl_set_pins", "eigkl_get_sizes",
    "eigkl_invalidate", "eigkl_get_stream", "eigkl_row_partition",
    "eigkl_assemble_laplacian", "eigkl_fiedler", "eigkl_partition_from_fiedler", "eigkl_write_eig",
    "eigkl_assemble_kl_graph", "eigkl_set_partition", "eigkl_set_partition_ordered", "eigkl_load_eig",
    "eigkl_kl_run", "eigkl_write_trace", "eigkl_get_partition", "eigkl_kl_rollback", "eigkl_write_partition", "eigkl_spmv", "eigkl_dvalues", "eigkl_cut",
    "eigkl_get_kl_values", "eigkl_get_node_order",
    "eigkl_get_laplacian", "eigkl_get_kl_graph", "eigkl_time_kernel",
]


class EigklError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code
        self.message = msg


class Opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("ncv", C.c_int32), ("max_restarts", C.c_int32), ("tol", C.c_double),
                ("keep", C.c_int32), ("seed", C.c_uint64), ("kl_cluster", C.c_int32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("struct_size", C.c_uint32),
                ("n_nodes", C.c_int64), ("n_nets", C.c_int64), ("n_pins", C.c_int64), ("n_pairs", C.c_int64),
                ("nnz_laplacian", C.c_int64), ("nnz_kl", C.c_int64),
                ("ncv", C.c_int32), ("matvecs", C.c_int32), ("restarts", C.c_int32), ("converged", C.c_int32),
                ("resid_est", C.c_double * 2), ("lambda_", C.c_double * 2),
                ("cheb_degree", C.c_int32), ("lanczos_steps", C.c_int32),
                ("kl_swaps", C.c_int64), ("kl_cluster", C.c_int32), ("kl_threads", C.c_int32),
                ("gpu_launches", C.c_int64),
                ("ms_assemble_laplacian", C.c_double), ("ms_assemble_kl", C.c_double), ("ms_fiedler", C.c_double),
                ("ms_partition", C.c_double), ("ms_kl_setup", C.c_double), ("ms_kl_loop", C.c_double),
                ("ms_spmv", C.c_double), ("ms_multidot", C.c_double), ("ms_update", C.c_double),
                ("ms_restart", C.c_double), ("ms_dvalues", C.c_double),
                ("n_spmv", C.c_int64), ("n_multidot", C.c_int64), ("n_update", C.c_int64), ("n_restart", C.c_int64),
                ("n_dvalues", C.c_int64),
                ("bytes_spmv", C.c_double), ("bytes_dvalues", C.c_double),
                ("bytes_multidot_total", C.c_double), ("bytes_update_total", C.c_double),
                ("spmv_per_launch", C.c_int32), ("resident_k", C.c_int32),
                ("gs_fused", C.c_int32), ("gs_cache_cols", C.c_int32),
                ("kl_local", C.c_int32), ("kl_flat", C.c_int32),
                ("dist_ranks", C.c_int32), ("dist_rows", C.c_int32), ("dist_halo", C.c_int64), ("dist_exports", C.c_int64),
                ("ms_comm", C.c_double), ("ms_push", C.c_double), ("n_comm", C.c_int64), ("n_push", C.c_int64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name.rstrip("_")] = list(v) if hasattr(v, "__len__") else v
        return d


class Trace(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("swaps", C.c_int64), ("cut", C.POINTER(C.c_float)),
                ("gain", C.POINTER(C.c_float)), ("node1", C.POINTER(C.c_int32)), ("node2", C.POINTER(C.c_int32))]


_lib = None


def load_library(path=LIB_PATH):
    """dlopen libeigkl.so and declare the prototypes.  Works without a GPU (no compute is run)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise OSError(f"{path} is missing: build it with `python -m eig_kl_algorithm_b200.build` "
                      "(there is no CPU fallback)")
    L = C.CDLL(path)
    P = C.POINTER
    H = C.c_void_p
    L.eigkl_abi_version.restype = C.c_int
    L.eigkl_nccl_unique_id.argtypes = [C.c_void_p]
    L.eigkl_create.argtypes = [P(H), P(Opts)]
    L.eigkl_destroy.argtypes = [H]
    L.eigkl_destroy.restype = None
    L.eigkl_last_error.argtypes = [H]
    L.eigkl_last_error.restype = C.c_char_p
    L.eigkl_get_stats.argtypes = [H, P(Stats)]
    L.eigkl_synchronize.argtypes = [H]
    L.eigkl_set_profile.argtypes = [H, C.c_int]
    L.eigkl_load_hgr.argtypes = [H, C.c_char_p]
    L.eigkl_set_pins.argtypes = [H, C.c_int32, C.c_int32, P(C.c_int64), P(C.c_int32)]
    L.eigkl_get_sizes.argtypes = [H, P(C.c_int32), P(C.c_int32), P(C.c_int64)]
    L.eigkl_invalidate.argtypes = [H]
    L.eigkl_get_stream.argtypes = [H, P(C.c_void_p)]
    L.eigkl_row_partition.argtypes = [C.c_int32, C.c_int32, C.c_int32, P(C.c_int32), P(C.c_int32), P(C.c_int32)]
    L.eigkl_assemble_laplacian.argtypes = [H]
    L.eigkl_fiedler.argtypes = [H, P(C.c_double), P(C.c_double)]
    L.eigkl_partition_from_fiedler.argtypes = [H, P(C.c_double), P(C.c_uint8)]
    L.eigkl_write_eig.argtypes = [H, C.c_char_p]
    L.eigkl_assemble_kl_graph.argtypes = [H]
    L.eigkl_set_partition.argtypes = [H, P(C.c_uint8)]
    L.eigkl_set_partition_ordered.argtypes = [H, P(C.c_int32), C.c_int64, P(C.c_int32), C.c_int64]
    L.eigkl_load_eig.argtypes = [H, C.c_char_p]
    L.eigkl_kl_run.argtypes = [H, P(Trace)]
    L.eigkl_write_trace.argtypes = [C.c_char_p, P(Trace)]
    L.eigkl_get_partition.argtypes = [H, P(C.c_uint8)]
    L.eigkl_kl_rollback.argtypes = [H, P(C.c_int64), P(C.c_float)]
    L.eigkl_write_partition.argtypes = [H, C.c_char_p]
    L.eigkl_spmv.argtypes = [H, P(C.c_double), P(C.c_double)]
    L.eigkl_dvalues.argtypes = [H, P(C.c_float)]
    L.eigkl_cut.argtypes = [H, P(C.c_float)]
    L.eigkl_get_kl_values.argtypes = [H, P(C.c_float)]
    L.eigkl_get_laplacian.argtypes = [H, P(C.c_int32), P(C.c_int32), P(C.c_double)]
    L.eigkl_get_node_order.argtypes = [H, P(C.c_int32)]
    L.eigkl_get_kl_graph.argtypes = [H, P(C.c_int32), P(C.c_int32), P(C.c_int32), P(C.c_float)]
    L.eigkl_time_kernel.argtypes = [H, C.c_int, C.c_int, C.c_int, P(C.c_double)]
    for s in SYMBOLS:
        f = getattr(L, s)
        if f.restype is C.c_int and s not in ("eigkl_abi_version",):
            f.restype = C.c_int
    _lib = L
    return L


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def row_partition(n_rows, nranks, rank):
    """(row_lo, row_hi, rows_padded) of the multi-rank row partition (host-only, no GPU needed)."""
    lo, hi, pad = C.c_int32(), C.c_int32(), C.c_int32()
    rc = load_library().eigkl_row_partition(n_rows, nranks, rank, C.byref(lo), C.byref(hi), C.byref(pad))
    if rc != 0:
        raise EigklError(rc, "eigkl_row_partition: bad arguments")
    return lo.value, hi.value, pad.value


def nccl_unique_id():
    buf = (C.c_ubyte * 128)()
    rc = load_library().eigkl_nccl_unique_id(buf)
    if rc != 0:
        raise EigklError(rc, load_library().eigkl_last_error(None).decode())
    return bytes(buf)


class Handle:
    """One GPU-resident EIG+KL problem (one handle per GPU / rank)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None, ncv=0, max_restarts=0, tol=0.0, keep=0, seed=0,
                 kl_cluster=0, flags=0):
        self.lib = load_library()
        self._h = C.c_void_p()
        o = Opts()
        o.struct_size = C.sizeof(Opts)
        o.device, o.rank, o.nranks = device, rank, nranks
        self._id = None
        if nccl_id is not None:
            self._id = (C.c_ubyte * 128).from_buffer_copy(nccl_id)
            o.nccl_unique_id = C.cast(self._id, C.c_void_p)
        o.ncv, o.max_restarts, o.tol, o.keep, o.seed = ncv, max_restarts, tol, keep, seed
        o.kl_cluster, o.flags = kl_cluster, flags
        rc = self.lib.eigkl_create(C.byref(self._h), C.byref(o))
        if rc != 0:
            raise EigklError(rc, self.lib.eigkl_last_error(None).decode())
        self.n_nodes = self.n_nets = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.eigkl_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise EigklError(rc, self.lib.eigkl_last_error(self._h).decode())

    # ---- input -----------------------------------------------------------------------------------
    def load_hgr(self, path):
        self._check(self.lib.eigkl_load_hgr(self._h, os.fsencode(path)))
        self._sizes()

    def set_pins(self, n_nodes, net_off, pins):
        net_off = np.ascontiguousarray(net_off, dtype=np.int64)
        pins = np.ascontiguousarray(pins, dtype=np.int32)
        self._check(self.lib.eigkl_set_pins(self._h, n_nodes, len(net_off) - 1, _ptr(net_off, C.c_int64), _ptr(pins, C.c_int32)))
        self._sizes()

    def _sizes(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self.lib.eigkl_get_sizes(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.n_nodes, self.n_nets, self.n_pins = a.value, b.value, c.value

    def set_pins_ptr(self, n_nodes, n_nets, net_off_ptr, pins_ptr):
        """Same as set_pins with raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        self._check(self.lib.eigkl_set_pins(self._h, n_nodes, n_nets, C.cast(net_off_ptr, C.POINTER(C.c_int64)),
                                            C.cast(pins_ptr, C.POINTER(C.c_int32))))
        self._sizes()

    def invalidate(self):
        self._check(self.lib.eigkl_invalidate(self._h))

    def stream_ptr(self):
        s = C.c_void_p()
        self._check(self.lib.eigkl_get_stream(self._h, C.byref(s)))
        return s.value or 0

    # ---- EIG ----------------------------------------------------------------------------------------
    def assemble_laplacian(self):
        self._check(self.lib.eigkl_assemble_laplacian(self._h))

    def fiedler(self, want_vector=True):
        lam = C.c_double()
        vec = np.empty(self.n_nodes, np.float64) if want_vector else None
        self._check(self.lib.eigkl_fiedler(self._h, C.byref(lam), _ptr(vec, C.c_double)))
        return lam.value, vec

    def partition_from_fiedler(self, want_side=True):
        med = C.c_double()
        side = np.empty(self.n_nodes, np.uint8) if want_side else None
        self._check(self.lib.eigkl_partition_from_fiedler(self._h, C.byref(med), _ptr(side, C.c_uint8)))
        return med.value, side

    def write_eig(self, path):
        self._check(self.lib.eigkl_write_eig(self._h, os.fsencode(path)))

    # ---- KL -------------------------------------------------------------------------------------------
    def assemble_kl_graph(self):
        self._check(self.lib.eigkl_assemble_kl_graph(self._h))

    def set_partition(self, side):
        side = np.ascontiguousarray(side, dtype=np.uint8)
        assert len(side) == self.n_nodes
        self._check(self.lib.eigkl_set_partition(self._h, _ptr(side, C.c_uint8)))

    def set_partition_ordered(self, order0, order1):
        o0 = np.ascontiguousarray(order0, dtype=np.int32)
        o1 = np.ascontiguousarray(order1, dtype=np.int32)
        self._check(self.lib.eigkl_set_partition_ordered(self._h, _ptr(o0, C.c_int32), len(o0), _ptr(o1, C.c_int32), len(o1)))

    def load_eig(self, path):
        self._check(self.lib.eigkl_load_eig(self._h, os.fsencode(path)))

    def kl_run(self, want_trace=True):
        if not want_trace:
            self._check(self.lib.eigkl_kl_run(self._h, None))
            return None
        cap = self.n_nodes // 2 + 2          # >= min(|left|, |right|) + 1 whatever the partition is
        cut = np.zeros(cap, np.float32)
        gain = np.zeros(cap, np.float32)
        n1 = np.zeros(cap, np.int32)
        n2 = np.zeros(cap, np.int32)
        t = Trace(cap, 0, _ptr(cut, C.c_float), _ptr(gain, C.c_float), _ptr(n1, C.c_int32), _ptr(n2, C.c_int32))
        self._check(self.lib.eigkl_kl_run(self._h, C.byref(t)))
        s = int(t.swaps) + 1
        return dict(swaps=int(t.swaps), cut=cut[:s], gain=gain[:s], node1=n1[:s], node2=n2[:s])

    def get_partition(self):
        side = np.empty(self.n_nodes, np.uint8)
        self._check(self.lib.eigkl_get_partition(self._h, _ptr(side, C.c_uint8)))
        return side

    def kl_rollback(self):
        """Undo the swaps after the best prefix of the last pass; returns (kept row, its cut)."""
        row, cut = C.c_int64(), C.c_float()
        self._check(self.lib.eigkl_kl_rollback(self._h, C.byref(row), C.byref(cut)))
        return row.value, np.float32(cut.value)

    def write_partition(self, path):
        self._check(self.lib.eigkl_write_partition(self._h, os.fsencode(path)))

    # ---- hooks ----------------------------------------------------------------------------------------
    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._check(self.lib.eigkl_spmv(self._h, _ptr(x, C.c_double), _ptr(y, C.c_double)))
        return y

    def dvalues(self):
        v = np.empty(self.n_nodes, np.float32)
        self._check(self.lib.eigkl_dvalues(self._h, _ptr(v, C.c_float)))
        return v

    def kl_values(self):
        v = np.empty(self.n_nodes, np.float32)
        self._check(self.lib.eigkl_get_kl_values(self._h, _ptr(v, C.c_float)))
        return v

    def cut(self):
        c = C.c_float()
        self._check(self.lib.eigkl_cut(self._h, C.byref(c)))
        return np.float32(c.value)

    def stats(self):
        s = Stats()
        s.struct_size = C.sizeof(Stats)
        self._check(self.lib.eigkl_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def get_laplacian(self):
        st = self.stats()
        n, nnz = self.n_nodes, st["nnz_laplacian"]
        rp = np.empty(n + 1, np.int32)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        self._check(self.lib.eigkl_get_laplacian(self._h, _ptr(rp, C.c_int32), _ptr(col, C.c_int32), _ptr(val, C.c_double)))
        return rp, col, val

    def node_order(self):
        perm = np.empty(self.n_nodes, np.int32)
        self._check(self.lib.eigkl_get_node_order(self._h, _ptr(perm, C.c_int32)))
        return perm

    def get_kl_graph(self):
        st = self.stats()
        n, nnz = self.n_nodes, st["nnz_kl"]
        rp = np.empty(n + 1, np.int32)
        fe = np.empty(n, np.int32)
        col = np.empty(max(nnz, 1), np.int32)[:nnz]
        w = np.empty(max(nnz, 1), np.float32)[:nnz]
        self._check(self.lib.eigkl_get_kl_graph(self._h, _ptr(rp, C.c_int32), _ptr(fe, C.c_int32), _ptr(col, C.c_int32), _ptr(w, C.c_float)))
        return rp, fe, col, w

    def time_kernel(self, what, iters=20, flush_l2=False):
        ms = C.c_double()
        self._check(self.lib.eigkl_time_kernel(self._h, {"spmv": 0, "dvalues": 1}[what], iters, int(flush_l2), C.byref(ms)))
        return ms.value

    def set_profile(self, on):
        self._check(self.lib.eigkl_set_profile(self._h, int(bool(on))))

    def synchronize(self):
        self._check(self.lib.eigkl_synchronize(self._h))


def write_trace(path, trace):
    cut = np.ascontiguousarray(trace["cut"], dtype=np.float32)
    gain = np.ascontiguousarray(trace["gain"], dtype=np.float32)
    t = Trace(len(cut), len(cut) - 1, _ptr(cut, C.c_float), _ptr(gain, C.c_float), None, None)
    rc = load_library().eigkl_write_trace(os.fsencode(path), C.byref(t))
    if rc != 0:
        raise EigklError(rc, load_library().eigkl_last_error(None).decode())


# ---------------------------------------------------------------------------------------------------
# Python mirrors of the reference's executables (same files, same names) -- used by the parity tests
# ---------------------------------------------------------------------------------------------------
def ceig(input_path, workdir=".", **opts):
    """cEIG <input>: writes <workdir>/pre_saved_EIG/<base>_out.txt (cEIG.cpp:138-237)."""
    os.makedirs(os.path.join(workdir, "results"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "pre_saved_EIG"), exist_ok=True)
    out = os.path.join(workdir, "pre_saved_EIG", os.path.basename(input_path) + "_out.txt")
    with Handle(**opts) as h:
        h.load_hgr(input_path)
        h.assemble_laplacian()
        lam, vec = h.fiedler()
        h.write_eig(out)
        st = h.stats()
    return dict(path=out, lambda2=lam, vec=vec, stats=st)


def ckl(input_path, eig=True, workdir=".", **opts):
    """cKL <input> -EIG: reads pre_saved_EIG/<base>_out.txt, writes results/<base>_KL_CutSize_EIG_output.txt."""
    os.makedirs(os.path.join(workdir, "results"), exist_ok=True)
    base = os.path.basename(input_path)
    assert eig, "the random branch needs an explicit order: use Handle.set_partition_ordered"
    out = os.path.join(workdir, "results", base + "_KL_CutSize_EIG_output.txt")
    with Handle(**opts) as h:
        h.load_hgr(input_path)
        h.assemble_kl_graph()
        h.load_eig(os.path.join(workdir, "pre_saved_EIG", base + "_out.txt"))
        tr = h.kl_run()
        write_trace(out, tr)
        st = h.stats()
        side = h.get_partition()
    return dict(path=out, trace=tr, stats=st, side=side)
