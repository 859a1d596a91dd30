#!/bin/bash
# Whole-step engine on the GPU box: the GPU test suite, then per-step timing against the
# other engines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python tools/step_bench.py --grids 1448,2048 --steps 100 --engines whole-step,persistent > gpurun_out/step_bench.jsonl 2> gpurun_out/step_bench.err
echo "step_bench exit $?"; cut -c1-150 gpurun_out/step_bench.jsonl; tail -3 gpurun_out/step_bench.err
