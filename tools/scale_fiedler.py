"""Row-partitioned Fiedler solve (and optionally the KL pass) of one circuit on WORLD_SIZE GPUs (torchrun).
usage: torchrun ... tools/scale_fiedler.py <workload> [kl]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from eig_kl_algorithm_b200 import api, datasets

name = sys.argv[1]
with_kl = len(sys.argv) > 2 and sys.argv[2] == "kl"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wd = "/tmp/eigkl_scale"
path = os.path.join(wd, "circuit", name + ".hgr")
if rank == 0:
    if name.startswith("synth"):
        datasets.write_synthetic(path, float(name[5:]))
    else:
        datasets.materialize(wd, circuits=(name,))
if world > 1:
    dist.barrier()
nccl_id = None
if world > 1:
    ids = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nccl_id = ids[0]
h = api.Handle(device=local, rank=rank, nranks=world, nccl_id=nccl_id)
h.load_hgr(path)
h.assemble_laplacian()
best = None
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lam, _ = h.fiedler(want_vector=False)
    st = h.stats()
    t = torch.tensor([st["ms_fiedler"]], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = float(t.item()) if best is None else min(best, float(t.item()))
out = {"workload": name, "n_gpus": world, "fiedler_ms": best, "spmv": st["matvecs"], "steps": st["lanczos_steps"], "lambda2": lam,
       "true_residual": st["resid_est"][1]}
if with_kl:
    h.partition_from_fiedler(want_side=False)
    h.assemble_kl_graph()
    h.kl_run(want_trace=False)
    st = h.stats()
    out.update(kl_loop_ms=st["ms_kl_loop"], kl_swaps=st["kl_swaps"])
if rank == 0:
    print("SCALE " + json.dumps(out), flush=True)
h.close()
if world > 1:
    dist.destroy_process_group()
