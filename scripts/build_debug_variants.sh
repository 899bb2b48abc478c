#!/usr/bin/env bash
# Timing experiments: the pipelined strip kernel with its slide or its gather compiled out
# (results are wrong by construction; only timings are of interest).  TIMING=1 adds the phase
# counters read by scripts/strip_timing.py.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/../pyparrm_b200/csrc" && pwd)"
root="$(cd "$here/../.." && pwd)"
mkdir -p "$root/build"
extra=""
[ "${TIMING:-0}" = "1" ] && extra="-DPARRM_STRIP_TIMING"
for variant in SKIP_SLIDE SKIP_GATHER; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
    -DPARRM_DEBUG_$variant $extra -I"$root/include" -I"$here" -shared -cudart static \
    "$here"/cabi.cu "$here"/taps.cu "$here"/filter.cu "$here"/filter_plan.cu "$here"/standardise.cu \
    "$here"/period_eval.cu -o "$root/build/libparrm_b200_$variant.so" 2>/dev/null &
done
# evaluator: phase counters of the tensor kernel (scripts/time_eval.py prints them when
# PARRM_TIMING_LIB points at this build), and the kernel with one of its phases compiled out
for variant in TENSOR_TIMING DEBUG_TENSOR_NO_MMA DEBUG_TENSOR_NO_GEN; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
    -DPARRM_$variant -I"$root/include" -I"$here" -shared -cudart static \
    "$here"/cabi.cu "$here"/taps.cu "$here"/filter.cu "$here"/filter_plan.cu "$here"/standardise.cu \
    "$here"/period_eval.cu -o "$root/build/libparrm_b200_$variant.so" 2>/dev/null &
done
wait
echo built debug variants
