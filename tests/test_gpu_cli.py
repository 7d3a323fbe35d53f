"""The three drop-in executables: argv, exit codes, file names and file contents."""
import os
import subprocess

import pytest

from conftest import GOLDEN
from eig_kl_algorithm_b200 import api

pytestmark = pytest.mark.gpu
BIN = api.BIN_DIR


def run(exe, args, cwd):
    return subprocess.run([os.path.join(BIN, exe)] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_ckl_eig_trace_file(eigkl_lib, workdir):
    for c in ("fract", "ibm01"):
        r = run("cKL", [f"circuit/{c}.hgr", "-EIG"], workdir)
        assert r.returncode == 0, r.stderr
        got = open(os.path.join(workdir, "results", f"{c}.hgr_KL_CutSize_EIG_output.txt"), "rb").read()
        assert got == open(os.path.join(GOLDEN, c + ".kl_trace_1core.txt"), "rb").read()


def test_gkl_same_engine(eigkl_lib, workdir):
    r = run("gKL", ["circuit/fract.hgr", "-EIG"], workdir)
    assert r.returncode == 0, r.stderr
    got = open(os.path.join(workdir, "results", "fract.hgr_KL_CutSize_EIG_output.txt"), "rb").read()
    assert got == open(os.path.join(GOLDEN, "fract.kl_trace_1core.txt"), "rb").read()


def test_ceig_then_ckl(eigkl_lib, workdir, tmp_path, oracle):
    import shutil
    import numpy as np
    wd = str(tmp_path)
    os.makedirs(os.path.join(wd, "circuit"))
    shutil.copy(os.path.join(workdir, "circuit", "fract.hgr"), os.path.join(wd, "circuit", "fract.hgr"))
    r = run("cEIG", ["circuit/fract.hgr"], wd)
    assert r.returncode == 0, r.stderr
    out = os.path.join(wd, "pre_saved_EIG", "fract.hgr_out.txt")
    assert os.path.isdir(os.path.join(wd, "results"))
    mine = oracle.read_eig(out, 149)
    gold = oracle.read_eig(os.path.join(workdir, "pre_saved_EIG", "fract.hgr_out.txt"), 149)
    assert abs(mine["lambda2"] - gold["lambda2"]) <= 1e-8 * gold["lambda2"]
    s = np.sign(mine["vec"] @ gold["vec"])
    assert np.abs(s * mine["vec"] - gold["vec"]).max() < 1e-6
    r = run("cKL", ["circuit/fract.hgr", "-EIG"], wd)
    assert r.returncode == 0, r.stderr
    # cKL consumed the file cEIG wrote: its trace must equal the oracle's pass from that same partition
    # (the eigenvector's sign is arbitrary, and for odd N a flipped sign is not an exact label swap, so
    # the golden trace only applies when the signs agree)
    got = open(os.path.join(wd, "results", "fract.hgr_KL_CutSize_EIG_output.txt"), "rb").read()
    o = oracle.OracleKL(oracle.OracleHgr(os.path.join(wd, "circuit", "fract.hgr")))
    ro = o.run(mine["side"])
    ref = str(tmp_path / "oracle_trace.txt")
    oracle.write_trace(ref, ro["cut"], ro["gain"])
    assert got == open(ref, "rb").read()
    if s > 0:
        assert got == open(os.path.join(GOLDEN, "fract.kl_trace_1core.txt"), "rb").read()


def test_cli_errors(eigkl_lib, workdir, tmp_path):
    r = run("cEIG", [], workdir)
    assert r.returncode == 1 and "Error: Usage: ./EIG <input_file>" in r.stderr          # cEIG.cpp:143-145,231-234
    r = run("cEIG", ["circuit/nope.hgr"], workdir)
    assert r.returncode == 1 and "Error opening input file" in r.stderr                  # cEIG.cpp:170-172
    r = run("cKL", [], workdir)
    assert r.returncode == 1 and "Usage:" in r.stdout                                    # cKL.cpp:431-434
    r = run("gKL", [], workdir)
    assert r.returncode == 1 and "Usage:" in r.stderr                                    # gKL.cu:673-676
    r = run("cKL", ["circuit/nope.hgr", "-EIG"], workdir)
    assert r.returncode == 1 and "Error opening file" in r.stderr                        # cKL.cpp:87-90
    wd = str(tmp_path)
    os.makedirs(os.path.join(wd, "circuit"))
    import shutil
    shutil.copy(os.path.join(workdir, "circuit", "fract.hgr"), os.path.join(wd, "circuit", "fract.hgr"))
    r = run("cKL", ["circuit/fract.hgr", "-EIG"], wd)
    assert r.returncode == 1 and "EIG file not found" in r.stderr                        # cKL.cpp:157-160
    r = run("cKL", ["circuit/fract.hgr"], wd)                                            # random branch runs
    assert r.returncode == 0 and os.path.exists(os.path.join(wd, "results", "fract.hgr_KL_CutSize_output.txt"))


def test_seed_rollback_and_partition_flags(eigkl_lib, workdir, tmp_path, oracle):
    """--seed makes the random branch (cKL.cpp:175-193) repeatable; --rollback keeps the best prefix
    of the pass (the reference tracks minCutSize, cKL.cpp:363, and stops there); --partition-out saves the sides."""
    import shutil
    import numpy as np
    wd = str(tmp_path)
    os.makedirs(os.path.join(wd, "circuit"))
    os.makedirs(os.path.join(wd, "pre_saved_EIG"))
    shutil.copy(os.path.join(workdir, "circuit", "ibm01.hgr"), os.path.join(wd, "circuit", "ibm01.hgr"))
    shutil.copy(os.path.join(workdir, "pre_saved_EIG", "ibm01.hgr_out.txt"), os.path.join(wd, "pre_saved_EIG", "ibm01.hgr_out.txt"))
    trace = os.path.join(wd, "results", "ibm01.hgr_KL_CutSize_output.txt")
    outs = []
    for seed, pout in (("7", "p7a.txt"), ("7", "p7b.txt"), ("8", "p8.txt")):
        r = run("cKL", ["circuit/ibm01.hgr", "--seed", seed, "--partition-out", pout], wd)
        assert r.returncode == 0, r.stderr
        outs.append((open(trace, "rb").read(), open(os.path.join(wd, pout), "rb").read()))
    assert outs[0] == outs[1]                    # same seed: identical trace and partition files
    assert outs[0][0] != outs[2][0]              # another seed: another start
    # --rollback from the golden start: the saved partition is the start with swaps 1..best applied
    r = run("cKL", ["circuit/ibm01.hgr", "-EIG", "--rollback"], wd)
    assert r.returncode == 0, r.stderr
    got = open(os.path.join(wd, "results", "ibm01.hgr_KL_CutSize_EIG_output.txt"), "rb").read()
    assert got == open(os.path.join(GOLDEN, "ibm01.kl_trace_1core.txt"), "rb").read()      # the trace itself is unchanged
    part = np.loadtxt(os.path.join(wd, "results", "ibm01.hgr_KL_partition_EIG.txt"), dtype=np.int64)
    g = oracle.read_eig(os.path.join(wd, "pre_saved_EIG", "ibm01.hgr_out.txt"), 12752)
    ro = oracle.OracleKL(oracle.OracleHgr(os.path.join(wd, "circuit", "ibm01.hgr"))).run(g["side"])
    best = int(np.argmin(ro["cut"]))             # first minimum
    expect = g["side"].copy()
    expect[ro["node1"][1:best + 1]] = 1
    expect[ro["node2"][1:best + 1]] = 0
    assert np.array_equal(part[:, 0], np.arange(12752)) and np.array_equal(part[:, 1].astype(np.uint8), expect)
    assert 0 < best < ro["swaps"] and int(expect.sum()) == int(g["side"].sum())
    # the reference's own argv forms are untouched by the extensions
    r = run("cKL", ["circuit/ibm01.hgr", "-EIG", "extra"], wd)
    assert r.returncode == 1 and "Usage:" in r.stdout
