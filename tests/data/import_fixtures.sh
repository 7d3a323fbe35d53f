#!/usr/bin/env bash
# Imports the reference's INPUT DATA (not code) as gzip fixtures so tests and bench.py can run on the
# GPU box, where /root/reference does not exist:
#   circuit/*.hgr              -> tests/data/circuit/<c>.hgr.gz          (hypergraph inputs)
#   pre_saved_EIG/*_out.txt    -> tests/data/pre_saved_EIG/<c>.hgr_out.txt.gz  (golden cEIG outputs)
# eig_kl_algorithm_b200.datasets.materialize() unpacks them into a working directory laid out
# like the reference's CWD (circuit/, pre_saved_EIG/, results/).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${EIGKL_REFERENCE_DIR:-/root/reference}
mkdir -p "$HERE/circuit" "$HERE/pre_saved_EIG"
for f in "$REF"/circuit/*.hgr; do gzip -9 -n -c "$f" > "$HERE/circuit/$(basename "$f").gz"; done
for f in "$REF"/pre_saved_EIG/*_out.txt; do gzip -9 -n -c "$f" > "$HERE/pre_saved_EIG/$(basename "$f").gz"; done
ls -la "$HERE/circuit" "$HERE/pre_saved_EIG"
