"""Row-partitioned SpMV chain timing under torchrun (tuning aid; EIGKL_DIST_DIAG experiments give invalid results).

    torchrun --nproc-per-node N tools/dist_diag.py <synth scale | circuit> [solve]

Prints, per rank, the per-SpMV time of a chain of row-partitioned SpMVs (halo pushed from the epilogue, flags awaited),
and with `solve` also one profiled Fiedler solve (per-class kernel times)."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("EIGKL_DIST", "rows")
from eig_kl_algorithm_b200 import api, datasets  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "synth10"
solve = len(sys.argv) > 2 and sys.argv[2] == "solve"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
wd = tempfile.mkdtemp()
if os.path.exists(name):
    path = name
elif name.startswith("synth"):
    path = datasets.write_synthetic(os.path.join(wd, name + ".hgr"), float(name[5:]))
else:
    path = datasets.materialize(wd, circuits=(name,))[name]
nid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    nid = ids[0]
h = api.Handle(device=local, rank=rank, nranks=world, nccl_id=nid, flags=api.EIGKL_F_PROFILE if solve else 0)
h.load_hgr(path)
h.assemble_laplacian()
st = h.stats()
warm = h.time_kernel("spmv", iters=64, flush_l2=False)
print(f"[rank {rank}/{world}] {name} diag={os.environ.get('EIGKL_DIST_DIAG', '0')} dist_ranks={st['dist_ranks']} rows={st['dist_rows']} "
      f"halo={st['dist_halo']} exports={st['dist_exports']}: chain {1e3 * warm:.2f} us per SpMV", flush=True)
if solve:
    lam, _ = h.fiedler(False)
    st = h.stats()
    per = lambda k: "%s %.2f us x%d" % (k, 1e3 * st["ms_" + k] / max(1, st["n_" + k]), st["n_" + k])
    print(f"[rank {rank}] fiedler {st['ms_fiedler']:.2f} ms lambda2 {lam:.3e} matvecs {st['matvecs']}: "
          + ", ".join(per(k) for k in ("spmv", "multidot", "update", "restart", "comm", "push")), flush=True)
h.close()
if world > 1:
    dist.destroy_process_group()
