#!/usr/bin/env bash
# Builds the UNMODIFIED reference KL program (and an instrumented twin that also prints the
# swapped node ids) from the sources where they lie under /root/reference.  Outputs go only to
# oracle/_ref/ (git-ignored, travels to the GPU box).  No reference source is copied into the repo:
# the instrumented twin is produced by piping a one-line sed edit straight into the compiler.
#
# cEIG.cpp is NOT buildable here: it needs Eigen3 + Spectra headers (cEIG.cpp:32-36), which are
# absent from the image and cannot be fetched (no network).  See DESIGN.md.
set -euo pipefail
REF=${EIGKL_REFERENCE_DIR:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF/cKL.cpp" ]; then
  echo "build_ref: $REF/cKL.cpp not present (GPU box?) - using prebuilt files in $OUT" >&2
  exit 0
fi
CXXFLAGS="-std=c++17 -O3 -fopenmp"          # Makefile:10-13 minus the unused conda include paths
g++ $CXXFLAGS "$REF/cKL.cpp" -o "$OUT/cKL"
# instrumented twin: trace row gains two columns (node1, node2); cKL.cpp:380
sed 's|fout << iteration << "\\t" << cutSize << "\\t" << gain << endl;|fout << iteration << "\\t" << cutSize << "\\t" << gain << "\\t" << node1 << "\\t" << node2 << endl;|' \
    "$REF/cKL.cpp" | g++ $CXXFLAGS -x c++ - -o "$OUT/cKL_instr"
# the reference's own GPU program, rebuilt for sm_100a (Makefile:14 with the arch changed); optional
if command -v nvcc >/dev/null 2>&1 && [ -f "$REF/gKL.cu" ]; then
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -use_fast_math -extended-lambda \
       -Wno-deprecated-declarations "$REF/gKL.cu" -o "$OUT/gKL_sm100a" 2>/dev/null || echo "build_ref: gKL build failed (optional)" >&2
fi
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
