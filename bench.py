#!/usr/bin/env python
"""bench.py -- the headline benchmark of the EIG+KL hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One step = one EIG+KL bipartition of the workload circuit: clique-Laplacian assembly (GPU sort +
segmented reduce) -> Fiedler solve (fp64 Lanczos) -> median split -> KL graph assembly -> one KL pass.
`value` is passes/s with the pins already resident in HBM; `e2e` is the same through the C ABI with
HOST buffers (pinned): H2D of the pins and D2H of the Fiedler vector, sides and KL trace every step.
Metric and workload follow BASELINE.json ("Fiedler solve ms & KL passes/sec on ibm18 ..."): ibm18.hgr
is a missing blob of the reference, so the largest shipped real circuit,
ibm10, stands in unless tests/data/circuit/ibm18.hgr(.gz) is present.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "KL passes/sec (1 pass = EIG+KL bipartition: Laplacian assembly + fp64 Lanczos Fiedler solve + KL pass)"
UNIT = "passes/s"


def pick_workload(name):
    data = os.path.join(ROOT, "tests", "data", "circuit")
    if name in (None, "auto"):
        for cand in ("ibm18",):
            if os.path.exists(os.path.join(data, cand + ".hgr")) or os.path.exists(os.path.join(data, cand + ".hgr.gz")):
                return cand
        return "ibm10"
    return name


def materialize_workload(name, workdir):
    from eig_kl_algorithm_b200 import datasets
    if name.startswith("synth"):
        scale = float(name[5:] or "1")
        return datasets.write_synthetic(os.path.join(workdir, "circuit", name + ".hgr"), scale), False
    plain = os.path.join(ROOT, "tests", "data", "circuit", name + ".hgr")
    if os.path.exists(plain):
        os.makedirs(os.path.join(workdir, "circuit"), exist_ok=True)
        dst = os.path.join(workdir, "circuit", name + ".hgr")
        shutil.copy(plain, dst)
        return dst, False
    paths = datasets.materialize(workdir, circuits=(name,), golden_eig=True)
    return paths[name], True


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places that may touch oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_port_pass(path, golden_side=None, max_restarts=0):
    """One EIG+KL pass of the oracle port (oracle/eigkl_oracle.c, OpenMP).  Returns seconds per stage."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as O
    t0 = time.perf_counter()
    oh = O.OracleHgr(path)
    t1 = time.perf_counter()
    e = O.OracleEIG(oh)
    t2 = time.perf_counter()
    lam, v, st = e.fiedler(max_restarts)
    t3 = time.perf_counter()
    side = (O.median(v) > v).astype(np.uint8)
    kl = O.OracleKL(oh)
    t4 = time.perf_counter()
    r = kl.run(side if golden_side is None else golden_side)
    t5 = time.perf_counter()
    return dict(parse=t1 - t0, assemble_l=t2 - t1, fiedler=t3 - t2, assemble_kl=t4 - t3, kl=t5 - t4,
                matvecs=st["matvecs"], converged=st["converged"], swaps=r["swaps"], lambda2=lam)


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores.
    KL = the UNMODIFIED reference program oracle/_ref/cKL (built from /root/reference/cKL.cpp by
    oracle/build_ref.sh).  EIG = the oracle port (cEIG.cpp needs Eigen+Spectra, absent from the image).
    Each step is a full pass of the workload; the number of steps is bounded by a time budget."""
    if rank != 0:
        return
    name = pick_workload(args.workload)
    wd = tempfile.mkdtemp(prefix="eigkl_ref_")
    path, has_golden = materialize_workload(name, wd)
    ckl = os.path.join(ROOT, "oracle", "_ref", "cKL")
    cores = host_threads()
    budget = float(os.environ.get("EIGKL_REF_BUDGET_S", "200"))
    use_ref_kl = os.path.exists(ckl) and has_golden
    times = []
    detail = {}
    t_begin = time.perf_counter()
    steps_done = 0
    warm = 0
    while steps_done < max(1, args.steps):
        t0 = time.perf_counter()
        p = cpu_port_pass(path)                           # EIG (port) + KL (port, used only if no reference binary)
        t_eig = p["parse"] + p["assemble_l"] + p["fiedler"]
        if use_ref_kl:
            t1 = time.perf_counter()
            r = subprocess.run([ckl, os.path.join("circuit", name + ".hgr"), "-EIG"], cwd=wd, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True)
            t_kl = time.perf_counter() - t1
            if r.returncode != 0:
                use_ref_kl = False
                t_kl = p["assemble_kl"] + p["kl"]
        else:
            t_kl = p["assemble_kl"] + p["kl"]
        dt = t_eig + t_kl
        detail = dict(eig_port_s=round(t_eig, 3), kl_s=round(t_kl, 3), matvecs=p["matvecs"], swaps=p["swaps"])
        if warm < args.warmup and (time.perf_counter() - t_begin) + 2 * dt < budget:
            warm += 1
            continue
        times.append(dt)
        steps_done += 1
        if (time.perf_counter() - t_begin) + dt > budget:
            break
    sec = sum(times) / len(times)
    val = 1.0 / sec
    kind = "reference" if use_ref_kl else "port"
    sample = (f"{len(times)} full pass(es) of {name}: Fiedler solve by the oracle port (cEIG unbuildable: Eigen/Spectra absent), "
              + ("KL by the unmodified reference cKL binary (oracle/_ref/cKL, all cores)" if use_ref_kl else "KL by the oracle port"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (Lanczos) + f32 (KL)", "data": "real circuit " + name + ".hgr" if not name.startswith("synth") else "synthetic",
            "config": {"workload": name, "requested_steps": args.steps, "requested_warmup": args.warmup, "time_budget_s": budget},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **detail},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", help="auto | fract | ibm01 | industry2 | ibm10 | ibm18 | synth<scale>")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kl-cluster", type=int, default=0)
    ap.add_argument("--keep", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from eig_kl_algorithm_b200 import api, build as _build, datasets

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the EIG+KL path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _build.build()
    name = pick_workload(args.workload)
    wd = tempfile.mkdtemp(prefix="eigkl_bench_%d_" % rank)
    path, _ = materialize_workload(name, wd)
    n_nodes, net_off, pins = datasets.read_hgr_arrays(path)
    n_nets = len(net_off) - 1
    # pinned host buffers: inputs and the results a caller reads back
    t_off = torch.from_numpy(net_off).pin_memory()
    t_pins = torch.from_numpy(pins).pin_memory()
    cap = n_nodes // 2 + 2
    t_vec = torch.empty(n_nodes, dtype=torch.float64).pin_memory()
    t_side = torch.empty(n_nodes, dtype=torch.uint8).pin_memory()
    t_cut = torch.empty(cap, dtype=torch.float32).pin_memory()
    t_gain = torch.empty(cap, dtype=torch.float32).pin_memory()
    t_n1 = torch.empty(cap, dtype=torch.int32).pin_memory()
    t_n2 = torch.empty(cap, dtype=torch.int32).pin_memory()

    import ctypes as C
    nccl_id = None
    if world > 1:
        ids = [api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    h = api.Handle(device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id, kl_cluster=args.kl_cluster, keep=args.keep)
    lib = h.lib
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))

    def upload():
        h.set_pins_ptr(n_nodes, n_nets, t_off.data_ptr(), t_pins.data_ptr())

    def step_resident():
        h.invalidate()
        h.assemble_laplacian()
        h.fiedler(want_vector=False)
        h.partition_from_fiedler(want_side=False)
        h.assemble_kl_graph()
        h.kl_run(want_trace=False)

    trace = api.Trace(cap, 0, C.cast(t_cut.data_ptr(), C.POINTER(C.c_float)), C.cast(t_gain.data_ptr(), C.POINTER(C.c_float)),
                      C.cast(t_n1.data_ptr(), C.POINTER(C.c_int32)), C.cast(t_n2.data_ptr(), C.POINTER(C.c_int32)))

    def step_e2e():
        upload()                                                                        # H2D: net_off + pins
        h.assemble_laplacian()
        lam = C.c_double()
        h._check(lib.eigkl_fiedler(h._h, C.byref(lam), C.cast(t_vec.data_ptr(), C.POINTER(C.c_double))))     # D2H: vector
        med = C.c_double()
        h._check(lib.eigkl_partition_from_fiedler(h._h, C.byref(med), C.cast(t_side.data_ptr(), C.POINTER(C.c_uint8))))  # D2H: sides
        h.assemble_kl_graph()
        h._check(lib.eigkl_kl_run(h._h, C.byref(trace)))                               # D2H: trace
        return lam.value, int(trace.swaps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(k):
            fn()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    upload()
    for _ in range(args.warmup):
        step_resident()
    l0 = h.stats()["gpu_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(step_resident, args.steps)
    launches = h.stats()["gpu_launches"] - l0
    st = h.stats()
    for _ in range(args.warmup):
        lam, swaps = step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()
    st_e2e = h.stats()
    h2d = int(net_off.nbytes + pins.nbytes)
    d2h = int(n_nodes * 8 + n_nodes + (swaps + 1) * 16 + 8 * 4)

    # N>1: ONE problem, row-partitioned Lanczos over the N ranks (strong scaling); the O(ms) assembly and
    # the latency-bound KL pass are replicated on every rank (multi-GPU)
    passes = args.steps
    value = passes / (ms_total * 1e-3)
    e2e_value = passes / (ms_e2e * 1e-3)

    # ---- per-kernel attribution: one extra profiled pass on the same data (events around every launch)
    roof, kernels = None, {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback, B200_PROFILING.md)"
    if rank == 0:
        with api.Handle(device=local_rank, kl_cluster=args.kl_cluster, keep=args.keep, flags=api.EIGKL_F_PROFILE) as hp:
            hp.set_pins_ptr(n_nodes, n_nets, t_off.data_ptr(), t_pins.data_ptr())
            for _ in range(2):
                hp.invalidate(); hp.assemble_laplacian(); hp.fiedler(want_vector=False); hp.partition_from_fiedler(want_side=False)
                hp.assemble_kl_graph(); hp.kl_run(want_trace=False)
            s0 = hp.stats()
            hp.invalidate(); hp.assemble_laplacian(); hp.fiedler(want_vector=False); hp.partition_from_fiedler(want_side=False)
            hp.assemble_kl_graph(); hp.kl_run(want_trace=False)
            s1 = hp.stats()
            spmv_iso = hp.time_kernel("spmv", iters=50, flush_l2=False)
            spmv_cold = hp.time_kernel("spmv", iters=20, flush_l2=True)
            dval_iso = hp.time_kernel("dvalues", iters=50, flush_l2=False)
            dval_cold = hp.time_kernel("dvalues", iters=20, flush_l2=True)

        def cls(nm, bytes_total=None, bytes_each=None):
            ms = s1["ms_" + nm] - s0["ms_" + nm]
            cnt = s1["n_" + nm] - s0["n_" + nm]
            if cnt <= 0:
                return None
            b = bytes_each * cnt if bytes_each is not None else bytes_total
            return {"launches": int(cnt), "ms_total": ms, "us_avg": 1e3 * ms / cnt, "alg_bytes_per_launch": b / cnt,
                    "achieved_gbs": b / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": b / (ms * 1e-3) / 1e9 / peak}
        kernels["spmv"] = cls("spmv", bytes_each=s1["bytes_spmv"])
        spl = int(s1.get("spmv_per_launch", 1) or 1)
        if kernels["spmv"] and spl > 1:
            # resident filter: ONE cooperative launch carries `spl` SpMVs (the library counts SpMVs); report per LAUNCH
            k = kernels["spmv"]
            k["spmv_count"] = k["launches"]
            k["launches"] = k["launches"] // spl
            k["spmv_per_launch"] = spl
            k["resident_k"] = int(s1.get("resident_k", 0))
            k["us_per_spmv"] = k["us_avg"]
            k["us_avg"] = k["us_avg"] * spl
            k["alg_bytes_per_launch"] = k["alg_bytes_per_launch"] * spl
        kernels["multidot"] = cls("multidot", bytes_total=s1["bytes_multidot_total"] - s0["bytes_multidot_total"])
        kernels["update"] = cls("update", bytes_total=s1["bytes_update_total"] - s0["bytes_update_total"])
        kernels["restart"] = cls("restart", bytes_total=(s1["n_restart"] - s0["n_restart"]) * (s1["ncv"] + max(3, s1["ncv"] // 5)) * n_nodes * 8.0)
        kernels["dvalues"] = cls("dvalues", bytes_each=s1["bytes_dvalues"])
        kernels["kl_loop"] = {"launches": 1, "ms_total": s1["ms_kl_loop"], "swaps": s1["kl_swaps"],
                              "us_per_swap": 1e3 * s1["ms_kl_loop"] / max(1, s1["kl_swaps"]),
                              "state": {0: "global memory (cluster kernel)", 1: "tile keys + side bits in shared memory",
                                        2: "tile keys in shared memory, state bytes in global memory"}.get(int(s1.get("kl_local", 0)), "?"),
                              "bound": "latency (a chain of dependent L2 round trips per swap), not bandwidth"}
        if kernels.get("multidot") and s1.get("gs_fused", 0):
            kernels["multidot"]["note"] = ("fused Gram-Schmidt: ONE cooperative launch per Lanczos step does both passes "
                                           "(h1, update, h2, update, DGKS decision, norm); %d basis columns cached in shared memory; "
                                           "algorithmic bytes = the basis once + w in and out" % int(s1.get("gs_cache_cols", 0)))
        kernels["spmv_isolated"] = {"us_avg_l2_warm": spmv_iso * 1e3, "us_avg_l2_flushed": spmv_cold * 1e3,
                                    "gbs_l2_warm": s1["bytes_spmv"] / (spmv_iso * 1e-3) / 1e9,
                                    "gbs_l2_flushed": s1["bytes_spmv"] / (spmv_cold * 1e-3) / 1e9}
        kernels["dvalues_isolated"] = {"us_avg_l2_warm": dval_iso * 1e3, "us_avg_l2_flushed": dval_cold * 1e3,
                                       "gbs_l2_warm": s1["bytes_dvalues"] / (dval_iso * 1e-3) / 1e9,
                                       "gbs_l2_flushed": s1["bytes_dvalues"] / (dval_cold * 1e-3) / 1e9}
        stream_classes = {k: v for k, v in kernels.items() if v and k in ("spmv", "multidot", "update", "restart", "dvalues")}
        dom = max(stream_classes, key=lambda k: stream_classes[k]["ms_total"])
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(name, {}).get(dom + "_resident" if dom == "spmv" and spl > 1 else dom)
        except Exception:
            pass
        d = stream_classes[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": d["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "us_avg": d["us_avg"], "launches_in_pass": d["launches"],
                "note": "average over every launch of this kernel in one extra profiled pass on the same resident data (CUDA events on the "
                        "library's stream). spmv: one launch of the resident filter kernel carries spmv_per_launch SpMVs (a whole Chebyshev "
                        "filter application; matrix in registers, x in shared memory, halo through L2), so its algorithmic bytes are "
                        "spmv_per_launch * (nnz*12 + n*20) while its DRAM traffic is one read of the matrix; when the matrix does not fit "
                        "on chip each SpMV is its own launch. The circuit is L2-resident: the fraction compares algorithmic bytes/time with "
                        "the HBM copy peak, it is not HBM utilisation (DESIGN.md section 4)"}
        if dom == "spmv" and spl > 1:
            roof["spmv_per_launch"] = spl

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        p = cpu_port_pass(path)
        sec = sum(p[k] for k in ("parse", "assemble_l", "fiedler", "assemble_kl", "kl"))
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": host_threads(), "kind": "port",
               "sample": f"one full pass of {name} by the oracle port (OpenMP C restatement of cEIG+cKL): "
                         f"{p['matvecs']} Lanczos matvecs, {p['swaps']} KL swaps",
               "seconds": {k: round(v, 4) for k, v in p.items() if isinstance(v, float) and k != "lambda2"}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "f64 (Lanczos) + f32 (KL, bit-exact with the reference)",
                "data": ("real circuit %s.hgr (ISPD98)" % name) if not name.startswith("synth") else "synthetic (seeded circuit_generator restatement)",
                "config": {"workload": name, "nodes": n_nodes, "nets": n_nets, "pins": int(len(pins)),
                           "parallelism": ("Lanczos row-partitioned over %d ranks (NCCL halo all-gather + dot all-reduces); KL D-values/arg-max partitioned by node range with one NCCL max all-reduce per swap; O(1 ms) assembly replicated" % world) if world > 1 else "1 GPU",
                           "l2": "working set < L2: every step re-assembles and re-solves from the resident pins; inputs are not flushed between steps",
                           "ncv": st["ncv"], "cheb_degree": st["cheb_degree"], "lanczos_steps": st["lanczos_steps"],
                           "spmv_per_pass": st["matvecs"], "restarts": st["restarts"],
                           "true_residual": st["resid_est"][1], "kl_swaps": st["kl_swaps"], "kl_cluster": st["kl_cluster"]},
                "stage_ms": {"assemble_laplacian": st["ms_assemble_laplacian"], "fiedler_solve": st["ms_fiedler"],
                             "partition": st["ms_partition"], "assemble_kl": st["ms_assemble_kl"], "kl_setup": st["ms_kl_setup"],
                             "kl_loop": st["ms_kl_loop"]},
                "fiedler_solve_ms": st["ms_fiedler"], "kl_pass_ms": st["ms_kl_setup"] + st["ms_kl_loop"],
                "kl_swaps_per_s": st["kl_swaps"] / max(1e-9, st["ms_kl_loop"] * 1e-3),
                "lambda2": lam, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "roofline": roof, "kernels": kernels, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
