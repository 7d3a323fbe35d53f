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
