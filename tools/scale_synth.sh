#!/bin/bash
# The 2 M-node synthetic circuit (BASELINE.json configs[4]) at N ranks: one bench line per N into gpurun_out/.
# usage (on a box with >= max(N) GPUs): bash tools/scale_synth.sh "1 2 4 8" [tag]
NS=${1:-"1 2"}; TAG=${2:-r02}
mkdir -p gpurun_out
for n in $NS; do
  out=gpurun_out/${TAG}_synth10_n$n.json
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --workload synth10 --extra '' --steps 3 --warmup 3 --no-legs --no-cpu-baseline > $out 2> gpurun_out/${TAG}_synth10_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) \
      bench.py --gpus $n --workload synth10 --extra '' --steps 3 --warmup 3 --no-legs --no-cpu-baseline > $out 2> gpurun_out/${TAG}_synth10_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("$out").read().strip().splitlines()[-1])
    k = d.get("kernels", {})
    print("N=$n value %.3f passes/s  fiedler %.1f ms  kl_loop %.1f ms  spmv %.1f us (frac %.2f)  parity %s" % (
        d["value"], d["fiedler_solve_ms"], d["stage_ms"]["kl_loop"], k["spmv"]["us_avg"], k["spmv"]["frac_of_hbm_peak"], (d.get("parity") or {}).get("ok")))
    ex = k.get("exchange")
    if ex: print("   exchange:", {a: ex[a] for a in ex if a != "note"})
except Exception as ex:
    print("N=$n failed:", ex); import subprocess; print(subprocess.run("tail -5 gpurun_out/${TAG}_synth10_n$n.err", shell=True, capture_output=True, text=True).stdout)
PY
done
