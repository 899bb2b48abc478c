#!/usr/bin/env bash
# compute-sanitizer over the GPU parity tests of the filter and the evaluator (memcheck, then
# racecheck on the shared-memory pipelines).  Output: gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
SEL='specialised_kernel_against_oracle[cfg2] or specialised_kernel_alignment or non_finite or golden_cases or device_nelder_mead or whole_golden_sweep or filter_sweep or compute_psd_matches'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 \
  python -m pytest tests/test_gpu_filter.py tests/test_gpu_search.py tests/test_gpu_taps.py tests/test_gpu_psd.py \
  -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 \
  python -m pytest tests/test_gpu_filter.py -m gpu -q -x -k "specialised_kernel_alignment or non_finite" \
  > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log
