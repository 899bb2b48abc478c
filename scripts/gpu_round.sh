#!/usr/bin/env bash
# One GPU visit: parity tests, filter shape sweep, bench.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if [ "${SWEEP:-1}" = "1" ]; then
  timeout 600 python scripts/sweep_filter.py > gpurun_out/sweep_filter.log 2>&1
  echo "sweep rc=$?"; tail -20 gpurun_out/sweep_filter.log
fi
if [ "${BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.log
fi
