"""The libstdc++ iteration-order replay against the REAL containers.

Three implementations must agree on every key sequence: std::unordered_map / unordered_set of this
toolchain (what the reference's fp32 summation order is defined by), the oracle's emulator
(oracle/eigkl_oracle.c) and the product's replay shared with the CUDA kernels (csrc/stl_order.h).
"""
import ctypes as C

import numpy as np
import pytest

from helpers import build_helpers


@pytest.fixture(scope="module")
def shim():
    L = C.CDLL(build_helpers.build())
    P = C.POINTER
    L.shim_stl_order.argtypes = [P(C.c_uint32), C.c_int32, P(C.c_int32)]
    L.shim_real_unordered_map_order.argtypes = [P(C.c_uint32), C.c_int32, P(C.c_uint32)]
    L.shim_real_unordered_set_order.argtypes = [P(C.c_uint32), C.c_int32, P(C.c_uint32)]
    L.shim_final_buckets.argtypes = [C.c_uint32]
    L.shim_final_buckets.restype = C.c_uint32
    L.shim_real_bucket_count.argtypes = [C.c_uint32]
    L.shim_real_bucket_count.restype = C.c_uint32
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _orders(shim, oracle, keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    n = len(keys)
    real_map = np.empty(n, np.uint32)
    real_set = np.empty(n, np.uint32)
    shim.shim_real_unordered_map_order(_p(keys, C.c_uint32), n, _p(real_map, C.c_uint32))
    shim.shim_real_unordered_set_order(_p(keys, C.c_uint32), n, _p(real_set, C.c_uint32))
    prod = np.empty(n, np.int32)
    shim.shim_stl_order(_p(keys, C.c_uint32), n, _p(prod, C.c_int32))
    orc = oracle.stl_hash_order(keys)
    return real_map, real_set, keys[prod], keys[orc]


def test_survey_example(shim, oracle):
    real_map, real_set, prod, orc = _orders(shim, oracle, [5, 18, 31, 2, 44])
    assert list(real_map) == [2, 44, 31, 18, 5]          # SURVEY.md Appendix E, measured
    assert list(prod) == [2, 44, 31, 18, 5] and list(orc) == [2, 44, 31, 18, 5]


@pytest.mark.parametrize("n", [0, 1, 2, 12, 13, 14, 28, 29, 30, 59, 60, 127, 128, 129, 541, 542, 1110, 5000])
def test_random_sequences(shim, oracle, n):
    rng = np.random.default_rng(1000 + n)
    for hi in (max(n, 1) * 2, 2_000_000):
        keys = rng.choice(hi, size=n, replace=False).astype(np.uint32)
        real_map, real_set, prod, orc = _orders(shim, oracle, keys)
        assert np.array_equal(real_map, real_set)        # GCC 13 range ctor = one-by-one inserts
        assert np.array_equal(prod, real_map)
        assert np.array_equal(orc, real_map)


def test_ascending_large_set(shim, oracle):
    # the shape calCutSize builds: right nodes in ascending id order (cKL.cpp:201)
    rng = np.random.default_rng(7)
    keys = np.sort(rng.choice(400_000, size=100_000, replace=False)).astype(np.uint32)
    real_map, real_set, prod, orc = _orders(shim, oracle, keys)
    assert np.array_equal(prod, real_set) and np.array_equal(orc, real_set)


def test_bucket_chain(shim):
    for n in [1, 13, 14, 29, 30, 59, 60, 127, 128, 257, 258, 541, 542, 1109, 1110, 2357, 2358, 5087, 5088, 20000, 100000]:
        assert shim.shim_final_buckets(n) == shim.shim_real_bucket_count(n), n
