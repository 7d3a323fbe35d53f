"""The N>1 host logic on CPU: world_size-2 gloo processes exercise the row partition and the
all-gather layout the multi-GPU Lanczos uses (csrc/internal.h:row_partition, lanczos.cu:lanczos_step),
with the CPU oracle standing in for the per-rank SpMV kernel."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from eig_kl_algorithm_b200 import api

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from eig_kl_algorithm_b200 import api
    import oracle_lib as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size={world})
    rank, world = dist.get_rank(), dist.get_world_size()
    h = O.OracleHgr({path!r})
    e = O.OracleEIG(h)
    n = e.n
    lo, hi, pad = api.row_partition(n, world, rank)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n)
    # every rank owns rows [lo, hi): computes its slice of y = L x, slices are gathered with equal,
    # padded counts -- global row g must land at index g of the gathered buffer
    y_full = e.spmv(x)
    mine = torch.zeros(pad, dtype=torch.float64)
    mine[: hi - lo] = torch.from_numpy(y_full[lo:hi].copy())
    got = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(got, mine)
    gathered = torch.cat(got).numpy()
    assert np.array_equal(gathered[:n], y_full)
    # dot products: local partial + all-reduce == global dot (what multidot + ncclAllReduce compute)
    part = torch.tensor([float(x[lo:hi] @ y_full[lo:hi])], dtype=torch.float64)
    dist.all_reduce(part)
    assert abs(part.item() - float(x @ y_full)) <= 1e-9 * abs(float(x @ y_full))
    # KL argmax exchange (C3): max over packed (orderable value : ~position) keys
    keys = torch.tensor([(int(rank + 1) << 32) | (0xFFFFFFFF - rank)], dtype=torch.int64)
    dist.all_reduce(keys, op=dist.ReduceOp.MAX)
    assert keys.item() == (world << 32) | (0xFFFFFFFF - (world - 1))
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_row_partition_properties(eigkl_lib):
    for n in (1, 31, 32, 33, 149, 12752, 69429, 2019200):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            pads = set()
            for r in range(world):
                lo, hi, pad = api.row_partition(n, world, r)
                assert lo == min(n, r * pad) and hi == min(n, (r + 1) * pad) and pad % 32 == 0
                seen += hi - lo
                pads.add(pad)
            assert seen == n and len(pads) == 1 and world * pads.pop() >= n


def test_two_rank_gloo_partition_and_gather(eigkl_lib, circuits, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=29731, world=2, path=circuits["ibm01"]))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
