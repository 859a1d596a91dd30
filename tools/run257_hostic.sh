#!/bin/bash
# The shipped 257^2 run with the initial fields built by the host loops (PDEODE_HOST_IC=1):
# the same start bits as the CPU oracle run, for the bitwise comparison of the early frames.
R=$GRAFT_REPO_ROOT
mkdir -p $R/gpurun_out/run257_hostic /tmp/run257h; cd /tmp/run257h
echo "$R/tests/data/shipped257.txt" | PDEODE_HOST_IC=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver > console.txt 2> err.txt
echo "exit $?"; grep "Time to compute" console.txt
cp total.dat statReport.dat console.txt $R/gpurun_out/run257_hostic/
python $R/tools/frames_to_npz.py output.dat $R/gpurun_out/run257_hostic/frames.npz 10
