"""In-tree build of libeigkl.so (hand-written sm_100a CUDA + C++17 host code) and the three CLIs.

    python -m eig_kl_algorithm_b200.build [--force] [--verbose]

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
    eig_kl_algorithm_b200/lib/libeigkl.so
    eig_kl_algorithm_b200/bin/{cEIG,cKL,gKL}
nvcc cross-compiles for sm_100a without a GPU.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
CLI = os.path.join(PKG, "cli")
OBJ = os.path.join(PKG, "build")
LIBDIR = os.path.join(PKG, "lib")
BINDIR = os.path.join(PKG, "bin")
LIB = os.path.join(LIBDIR, "libeigkl.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"          # the image exports CXX=/opt/gcc/bin/g++, which lacks libgomp.spec
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
                     "-ccbin", HOST_CXX, "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
                     "--expt-relaxed-constexpr", "-Xptxas", "-v"]
SOURCES = ["scan_sort.cu", "assemble.cu", "spmv.cu", "lanczos.cu", "dist.cu", "kl.cu", "dense_eig.cpp", "hgr_io.cpp",
           "cabi.cpp", "comm.cpp"]
HEADERS = ["internal.h", "device_utils.cuh", "stl_order.h", os.path.join(ROOT, "include", "eigkl.h")]
CLIS = ["cEIG", "cKL", "gKL"]
WITH_NCCL = os.path.exists("/usr/include/nccl.h") and os.environ.get("EIGKL_NO_NCCL") is None


def _stamp(paths, extra=""):
    hsh = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            hsh.update(f.read())
    return hsh.hexdigest()


def _run(cmd, verbose, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + cmd[-1])
    if verbose:
        sys.stdout.write(r.stdout)
    return r.stdout


def build(force=False, verbose=False):
    for d in (OBJ, LIBDIR, BINDIR):
        os.makedirs(d, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    flags = list(NVCC_FLAGS) + (["-DEIGKL_WITH_NCCL"] if WITH_NCCL else [])
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s + ".o")
        st = _stamp([src] + hdrs, " ".join(flags))
        stf = obj + ".stamp"
        if force or not os.path.exists(obj) or not os.path.exists(stf) or open(stf).read() != st:
            jobs.append((src, obj, st, stf))
    def compile_one(j):
        src, obj, st, stf = j
        extra = ["-x", "cu"] if src.endswith(".cpp") else []
        _run([NVCC] + flags + extra + ["-c", src, "-o", obj], verbose, log=obj + ".log")
        with open(stf, "w") as f:
            f.write(st)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s + ".o") for s in SOURCES]
    if jobs or force or not os.path.exists(LIB):
        link = [NVCC] + ARCH + ["-shared", "-ccbin", HOST_CXX, "-o", LIB] + objs + ["-lcudart"]
        link += ["-ldl"]          # NCCL is bound with dlopen at run time (see csrc/comm.cpp)
        _run(link, verbose)
    cli_hdrs = [os.path.join(CLI, "cli_common.h"), os.path.join(CLI, "kl_main.h"), os.path.join(ROOT, "include", "eigkl.h")]
    for c in CLIS:
        src = os.path.join(CLI, c + ".cpp")
        out = os.path.join(BINDIR, c)
        st = _stamp([src] + cli_hdrs)
        stf = os.path.join(OBJ, c + ".stamp")
        if force or jobs or not os.path.exists(out) or not os.path.exists(stf) or open(stf).read() != st:
            _run([HOST_CXX, "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", out,
                  "-L" + LIBDIR, "-leigkl", "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-rpath-link," + LIBDIR,
                  "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], verbose)
            with open(stf, "w") as f:
                f.write(st)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
