#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, bench (both arms), launch list of the bench command.
# Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
if [ "${BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench rc=$?"; tail -c 600 gpurun_out/bench.log
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
  echo "reference rc=$?"; tail -c 800 gpurun_out/bench_reference.log
fi
if [ "${LAUNCHES:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/bench_launch_list.csv python bench.py --steps 5 --warmup 3 \
    > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?"
fi
