#!/bin/bash
# wall-clock breakdown of the drop-in executables (EIGKL_TIMING=1), run on the GPU box
set -e
W=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, "."); from eig_kl_algorithm_b200 import datasets; datasets.materialize("$W", circuits=("ibm01","ibm10"))
PY
cd $W
for c in ibm01 ibm10; do
  echo "== cEIG $c"; EIGKL_TIMING=1 $OLDPWD/eig_kl_algorithm_b200/bin/cEIG circuit/$c.hgr 2>&1 >/dev/null | grep timing
  echo "== cKL $c";  EIGKL_TIMING=1 $OLDPWD/eig_kl_algorithm_b200/bin/cKL circuit/$c.hgr -EIG 2>&1 >/dev/null | grep timing
done
