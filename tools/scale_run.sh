#!/bin/bash
# 1/2/4/8-GPU runs of the multi-rank tests and of bench.py (what the driver does at round end); writes gpurun_out/
# usage (on a box with 8 GPUs): bash tools/scale_run.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi_tests.log 2>&1
tail -3 gpurun_out/multi_tests.log
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", "value", round(d["value"], 3), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3))
except Exception as ex:
    print("N=$n failed:", ex)
PY
done
