#!/bin/bash
# Small-grid engines on the GPU box: GPU test suite, time per step by engine, and the two
# shipped configurations end to end with the host loop's own timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python tools/step_bench.py --grids 257,512,1024 --steps 200 > gpurun_out/step_bench.jsonl 2> gpurun_out/step_bench.err
echo "step_bench exit $?"; cut -c1-150 gpurun_out/step_bench.jsonl; tail -3 gpurun_out/step_bench.err
R=$GRAFT_REPO_ROOT
mkdir -p /tmp/w1 && cd /tmp/w1
for f in shipped257.txt shipped257_wave.txt; do
  echo "$R/tests/data/$f" | PDEODE_HOST_TIMING=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver 2>&1 | grep "Time to compute\|host timing\|Avg Iters"
done
