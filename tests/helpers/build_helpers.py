"""Builds tests/helpers/libhostshim.so: the product's HOST code compiled with g++ for the CPU suite."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "eig_kl_algorithm_b200", "csrc")
OUT = os.path.join(HERE, "libhostshim.so")
SRCS = [os.path.join(HERE, "host_shim.cpp"), os.path.join(CSRC, "hgr_io.cpp"), os.path.join(CSRC, "dense_eig.cpp")]
DEPS = SRCS + [os.path.join(CSRC, h) for h in ("internal.h", "stl_order.h")]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-I" + CSRC,
                           "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"] + SRCS +
                          ["-o", OUT, "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return OUT
