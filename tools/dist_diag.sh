#!/bin/bash
# Row-partitioned SpMV chain under the timing experiments of EIGKL_DIST_DIAG (results invalid for diag != 0).
# usage (on a box with N GPUs): bash tools/dist_diag.sh N [scale]
N=${1:-2}; SCALE=${2:-10.0}
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0, '.')
from eig_kl_algorithm_b200 import datasets
datasets.write_synthetic('/tmp/synth_diag.hgr', float('$SCALE'))
"
for d in ${DIAGS:-0 1 2 3}; do
  EIGKL_DIST_DIAG=$d timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + d)) \
    tools/dist_diag.py /tmp/synth_diag.hgr $([ "$d" = "0" ] && echo solve) 2>&1 | grep -E "^\[rank" 
done
