#!/usr/bin/env bash
# Regenerates the KL known-answer fixtures by RUNNING THE REFERENCE ITSELF (oracle/_ref/cKL, built
# by oracle/build_ref.sh from /root/reference/cKL.cpp) on the shipped circuits with the shipped
# pre_saved_EIG partitions.  Only works in the build container (needs /root/reference).
#   <c>.kl_trace_1core.txt : unmodified trace file of `taskset -c 0 cKL circuit/<c>.hgr -EIG`
#                            (1 core => the float cut reduction of cKL.cpp:203 is deterministic)
#   <c>.kl_swaps.txt       : trace of the instrumented twin: iter, cut, gain, node1, node2
# usage: make_golden.sh [circuit ...]   (default: fract ibm01 industry2; ibm10 takes ~15 min on 1 core)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${EIGKL_REFERENCE_DIR:-/root/reference}
BIN="$HERE/../../oracle/_ref"
W=$(mktemp -d)
trap 'rm -rf "$W"' EXIT
cp -r "$REF/pre_saved_EIG" "$W/"
cd "$W"
CIRCUITS=("$@"); [ ${#CIRCUITS[@]} -eq 0 ] && CIRCUITS=(fract ibm01 industry2)
for c in "${CIRCUITS[@]}"; do
  taskset -c 0 "$BIN/cKL" "$REF/circuit/$c.hgr" -EIG > /dev/null
  cp "results/$c.hgr_KL_CutSize_EIG_output.txt" "$HERE/$c.kl_trace_1core.txt"
  "$BIN/cKL_instr" "$REF/circuit/$c.hgr" -EIG > /dev/null
  cp "results/$c.hgr_KL_CutSize_EIG_output.txt" "$HERE/$c.kl_swaps.txt"
  echo "$c full-md5=$(md5sum < "$HERE/$c.kl_trace_1core.txt" | cut -d' ' -f1) gaincol-md5=$(cut -f1,3 "$HERE/$c.kl_trace_1core.txt" | md5sum | cut -d' ' -f1) swapseq-md5=$(tail -n +2 "$HERE/$c.kl_swaps.txt" | cut -f4,5 | md5sum | cut -d' ' -f1)"
done
