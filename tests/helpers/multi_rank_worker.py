"""Worker for tests/test_gpu_multi.py: one process per GPU (torch.distributed.run).  Every rank solves
the same circuit with the row-partitioned Lanczos over NCCL; rank 0 compares with a single-GPU solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eig_kl_algorithm_b200 import api  # noqa: E402

path = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ids = [api.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
h = api.Handle(device=local, rank=rank, nranks=world, nccl_id=ids[0])
h.load_hgr(path)
h.assemble_laplacian()
x = np.random.default_rng(1).standard_normal(h.n_nodes)
y = h.spmv(x)                                   # row slices + all-gather
lam, v = h.fiedler()
med, side = h.partition_from_fiedler()
h.assemble_kl_graph()
tr = h.kl_run()
st = h.stats()
ok = True
if rank == 0:
    with api.Handle(device=local) as h1:
        h1.load_hgr(path)
        h1.assemble_laplacian()
        y1 = h1.spmv(x)
        lam1, v1 = h1.fiedler()
        med1, side1 = h1.partition_from_fiedler()
        h1.assemble_kl_graph()
        tr1 = h1.kl_run()
    cs = abs(v @ v1)
    sine = float(np.sqrt(max(0.0, 1.0 - cs * cs)))
    checks = {
        # same rows, but the row blocks (hence lanes per row / summation order) differ with the partition
        "spmv_equal": bool(np.abs(y - y1).max() <= 1e-13 * np.abs(y1).max()),
        "lambda_rel": abs(lam - lam1) / abs(lam1),
        "sine": sine,
        "sides_equal_up_to_sign": bool(np.array_equal(side, side1) or (side != side1).sum() in (0, 1, len(side), len(side) - 1)),
        "swaps": (tr["swaps"], tr1["swaps"]),
    }
    ok = checks["spmv_equal"] and checks["lambda_rel"] < 1e-9 and sine < 1e-7
    # KL across ranks (NCCL arg-max exchange) must reproduce the single-GPU pass bit for bit from the same partition
    with api.Handle(device=local) as h2:
        h2.load_hgr(path)
        h2.assemble_kl_graph()
        h2.set_partition(side)
        trs = h2.kl_run()
    checks["kl_multi_equals_single"] = bool(tr["swaps"] == trs["swaps"] and np.array_equal(tr["node1"], trs["node1"])
                                            and np.array_equal(tr["node2"], trs["node2"])
                                            and np.array_equal(tr["cut"].view(np.uint32), trs["cut"].view(np.uint32))
                                            and np.array_equal(tr["gain"].view(np.uint32), trs["gain"].view(np.uint32)))
    ok = ok and checks["kl_multi_equals_single"]
    print("MULTI", world, "lambda2", lam, "matvecs", st["matvecs"], "fiedler_ms %.2f" % st["ms_fiedler"], "kl_loop_ms %.2f" % st["ms_kl_loop"], checks, "OK" if ok else "FAIL", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
h.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
