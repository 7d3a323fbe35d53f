"""Multi-GPU path, one process per GPU.  Needs >= 2 GPUs on the box; skipped otherwise.

rows: row-partitioned Lanczos (nnz-balanced cuts, packed halos pushed over NVLink from the SpMV epilogue, NCCL
all-reduces for the dot products), forced even for circuits that fit one chip so that the path is exercised on the
reference's circuits -- industry2 brings the long rows, ibm10 the size, the synthetic circuit a disconnected graph.
auto: what the library does by default (chip-resident matrices are solved replicated on every rank)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except OSError:
        return 0


CASES = [(2, "ibm01", "rows"), (2, "industry2", "rows"), (2, "ibm10", "rows"), (2, "ibm10", "auto"), (2, "synth1", "rows"),
         (4, "ibm01", "rows"), (4, "ibm10", "auto"), (4, "synth1", "rows"),
         (8, "ibm01", "rows"), (8, "industry2", "rows"), (8, "ibm10", "auto"), (8, "synth1", "rows")]


@pytest.mark.parametrize("world,circuit,mode", CASES)
def test_multi_rank_matches_single_gpu(world, circuit, mode, eigkl_lib, circuits, tmp_path_factory):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    if circuit.startswith("synth"):
        from eig_kl_algorithm_b200 import datasets
        path = datasets.write_synthetic(os.path.join(str(tmp_path_factory.getbasetemp()), circuit + ".hgr"), float(circuit[5:]))
    else:
        path = circuits[circuit]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "helpers", "multi_rank_worker.py"), path, mode]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert " OK" in r.stdout


def test_partitioned_kl_option_matches_local_loop(eigkl_lib, circuits):
    """EIGKL_KL_DIST=1 keeps round 1's node-partitioned KL (one NCCL arg-max per swap) as a tested option."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, EIGKL_KL_DIST="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "helpers", "multi_rank_worker.py"), circuits["ibm01"], "rows"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-4000:]
    assert " OK" in r.stdout
