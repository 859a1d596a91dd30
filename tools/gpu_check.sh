#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, a short bench and,
# when all of that is green, the ncu launch list of a short bench run.
# Usage: tools/gpu_check.sh [tests|bench|ncu|all]
MODE=${1:-all}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
if [ "$MODE" = tests ] || [ "$MODE" = all ]; then
  echo "== pytest -m gpu"
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -n 30 gpurun_out/pytest_gpu.log
  echo "== smoke"
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
fi
if [ "$MODE" = bench ] || [ "$MODE" = all ]; then
  echo "== bench"
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -n 5 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [ "$MODE" = ncu ] || [ "$MODE" = all ]; then
  echo "== ncu launch list"
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --kernel-reps 5"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 \
      --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_run.log; wc -l gpurun_out/launches.csv
fi
