#!/bin/bash
mkdir -p gpurun_out
R=$GRAFT_REPO_ROOT
mkdir -p /tmp/w1 && cd /tmp/w1
echo "$R/tests/data/shipped257.txt" | PDEODE_HOST_TIMING=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver 2>&1 | grep "Time to compute\|host timing"
sed 's/^tEnd .*/tEnd      = 0.02/' $R/tests/data/shipped257.txt > short.txt
echo "short.txt" | PDEODE_STEP_STAMPS=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver 2>&1 | grep "step stamps" | tail -3 | cut -c1-400
