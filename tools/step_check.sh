#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
R=$GRAFT_REPO_ROOT
mkdir -p /tmp/w1 && cd /tmp/w1
for f in shipped257.txt shipped257_wave.txt; do
echo "$R/tests/data/$f" | PDEODE_HOST_TIMING=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver 2>&1 | grep "Time to compute\|host timing\|Avg Iters"
done
