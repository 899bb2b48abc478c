#!/usr/bin/env bash
# Build libparrm_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$here/../libparrm_b200.so"
obj="$here/build"
mkdir -p "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
       -Xcompiler -fPIC -I"$root/include" -I"$here" ${PARRM_NVCC_EXTRA:-})
pids=()
for src in cabi taps filter filter_plan standardise period_eval; do
  "$NVCC" "${FLAGS[@]}" -c "$here/$src.cu" -o "$obj/$src.o" &
  pids+=($!)
done
for pid in "${pids[@]}"; do wait "$pid"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out" "$obj"/cabi.o "$obj"/taps.o "$obj"/filter.o "$obj"/filter_plan.o "$obj"/standardise.o \
        "$obj"/period_eval.o -cudart static
echo "built $out"
