#!/bin/bash
# BASELINE config 2 on the GPU box: the shipped parameters on a 1024^2 grid, all 30 001 steps;
# every 10th frame, total.dat and the console come back for the comparison with the CPU oracle run.
R=$GRAFT_REPO_ROOT
mkdir -p $R/gpurun_out/run1024 /tmp/run1024; cd /tmp/run1024
echo "$R/tests/data/shipped1024.txt" | PDEODE_HOST_TIMING=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver > console.txt 2> err.txt
echo "exit $?"; tail -9 console.txt; cat err.txt | tail -2
ls -la
python $R/tools/frames_to_npz.py output.dat $R/gpurun_out/run1024/frames.npz 10
cp total.dat COprod.dat statReport.dat console.txt $R/gpurun_out/run1024/
