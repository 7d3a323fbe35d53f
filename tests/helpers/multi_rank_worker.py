"""Worker for tests/test_gpu_multi.py: one process per GPU (torch.distributed.run).

    multi_rank_worker.py <circuit.hgr> <rows|auto>

rows: the Lanczos solve is forced onto the row-partitioned path (EIGKL_DIST=rows: nnz-balanced row cuts, packed
halo slots in the peer-mapped arena, exports pushed from the SpMV epilogue) whatever the matrix size; auto: the
library decides (a matrix that fits one chip is solved replicated on every rank).  Rank 0 compares with a
single-GPU handle: SpMV, lambda2, sine of the Fiedler vectors (when lambda2 is simple), and the KL trace bit for
bit from the same partition."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

path, mode = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "rows")
if mode == "rows":
    os.environ["EIGKL_DIST"] = "rows"
from eig_kl_algorithm_b200 import api  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ids = [api.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
h = api.Handle(device=local, rank=rank, nranks=world, nccl_id=ids[0])
h.load_hgr(path)
h.assemble_laplacian()
x = np.random.default_rng(1).standard_normal(h.n_nodes)
y = h.spmv(x)                                   # row slices, halo pushed between the ranks, result gathered
lam, v = h.fiedler()
med, side = h.partition_from_fiedler()
h.assemble_kl_graph()
tr = h.kl_run()
st = h.stats()
# a second solve on the same handle (arena, flags and sequence numbers are reused)
h.invalidate(); h.assemble_laplacian()
lam_b, v_b = h.fiedler()
ok = True
if rank == 0:
    os.environ.pop("EIGKL_DIST", None)
    with api.Handle(device=local) as h1:
        h1.load_hgr(path)
        h1.assemble_laplacian()
        y1 = h1.spmv(x)
        lam1, v1 = h1.fiedler()
        st1 = h1.stats()
        med1, side1 = h1.partition_from_fiedler()
        h1.assemble_kl_graph()
        tr1 = h1.kl_run()
    cs = abs(v @ v1)
    sine = float(np.sqrt(max(0.0, 1.0 - cs * cs)))
    simple = abs(lam1) > 1e-9                   # lambda2 = 0: disconnected graph, the null space has no preferred vector
    checks = {
        "dist_ranks": st["dist_ranks"], "halo": st["dist_halo"], "exports": st["dist_exports"], "rows": st["dist_rows"],
        # same rows, but the row blocks (hence the summation order inside a row) differ with the partition
        "spmv_equal": bool(np.abs(y - y1).max() <= 1e-13 * np.abs(y1).max()),
        "lambda_err": abs(lam - lam1) / abs(lam1) if simple else abs(lam - lam1),
        "sine": sine if simple else None,
        "resid": st["resid_est"][1],
        "repeat_equal": bool(lam == lam_b and np.array_equal(v, v_b)),
        "swaps": (tr["swaps"], tr1["swaps"]),
    }
    ok = checks["spmv_equal"] and checks["repeat_equal"] and st["converged"] == 1
    ok = ok and checks["lambda_err"] < 1e-9       # relative when lambda2 is simple, absolute (eigenvalues are O(1)) when it is 0
    if simple:
        ok = ok and sine < 1e-7
    if mode == "rows":
        ok = ok and st["dist_ranks"] == world
    if st["dist_ranks"] == 1:                   # replicated: the very same kernels ran -> bit-identical results
        checks["replica_identical"] = bool(lam == lam1 and np.array_equal(v, v1) and np.array_equal(side, side1)
                                           and np.array_equal(tr["node1"], tr1["node1"]) and np.array_equal(tr["cut"].view(np.uint32), tr1["cut"].view(np.uint32)))
        ok = ok and checks["replica_identical"]
    # the KL pass of the multi-rank handle must reproduce a single-GPU pass from the same partition bit for bit
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
    print("MULTI", world, mode, os.path.basename(path), "lambda2", lam, "matvecs", st["matvecs"], "fiedler_ms %.2f (1 GPU %.2f)" % (st["ms_fiedler"], st1["ms_fiedler"]),
          "kl_loop_ms %.2f" % st["ms_kl_loop"], checks, "OK" if ok else "FAIL", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
h.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
