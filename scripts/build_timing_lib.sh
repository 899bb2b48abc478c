#!/usr/bin/env bash
# Debug build of the library with phase counters in the evaluator kernels
# (-DPARRM_SOLVE_TIMING and/or -DPARRM_TENSOR_TIMING), next to the product library:
#   bash scripts/build_timing_lib.sh -DPARRM_SOLVE_TIMING
#   PARRM_TIMING_LIB=scripts/_build/libparrm_timing.so EVAL_CANDIDATES=5 python scripts/time_eval.py
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="$root/pyparrm_b200/csrc"
out="$root/scripts/_build"
mkdir -p "$out"
bash "$src/build.sh" > /dev/null   # the other objects
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
     -I"$root/include" -I"$src" -I"$src/build" "$@" -c "$src/period_eval.cu" -o "$out/period_eval_timing.o"
objs=()
for name in cabi taps filter filter_plan filter_jit standardise psd neldermead host_copy; do
  objs+=("$src/build/$name.o")
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libparrm_timing.so" \
     "${objs[@]}" "$out/period_eval_timing.o" -cudart static -ldl
echo "built $out/libparrm_timing.so"
