#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python tools/step_bench.py --grids 1448,2048 --steps 100 --engines persistent 2>/dev/null | cut -c1-150
PDEODE_PERSIST=1 timeout 300 python bench.py --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null | cut -c1-200
