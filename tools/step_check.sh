#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "whole or bit_exact or long_run or cap or warm" > gpurun_out/pytest_step.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_step.log
timeout 300 python tools/step_bench.py --grids 257,512,1024 --steps 300 --groups 1,2,3 --engines whole-step 2>/dev/null | cut -c1-150
