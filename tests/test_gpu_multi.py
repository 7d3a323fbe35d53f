"""Multi-GPU path: row-partitioned Lanczos with NCCL halo all-gather and dot-product
all-reduces, one process per GPU.  Needs >= 2 GPUs on the box; skipped otherwise."""
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


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("circuit", ["ibm01"])
def test_row_partitioned_lanczos_matches_single_gpu(world, circuit, eigkl_lib, circuits):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "helpers", "multi_rank_worker.py"), circuits[circuit]]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "OK" in r.stdout
