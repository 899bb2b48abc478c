#!/usr/bin/env bash
# Debug build with per-phase cycle counters in the strip kernel -> build/libparrm_b200_timing.so
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/../pyparrm_b200/csrc" && pwd)"
root="$(cd "$here/../.." && pwd)"
mkdir -p "$root/build"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  -DPARRM_STRIP_TIMING -I"$root/include" -I"$here" -shared -cudart static \
  "$here"/cabi.cu "$here"/taps.cu "$here"/filter.cu "$here"/filter_plan.cu "$here"/standardise.cu \
  "$here"/period_eval.cu -o "$root/build/libparrm_b200_timing.so"
echo built "$root/build/libparrm_b200_timing.so"
