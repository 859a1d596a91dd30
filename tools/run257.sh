#!/bin/bash
# The reference's shipped configuration end to end on the GPU box:
# 257^2 (runAll.sh mode) and 257x4 (twoDimension_runall.sh mode), 30001 steps each.
# Every 10th frame, total.dat and the console come back for comparisons with CPU runs.
R=$GRAFT_REPO_ROOT
for tag in run257:shipped257.txt run257w:shipped257_wave.txt; do
  d=${tag%%:*}; f=${tag##*:}
  mkdir -p $R/gpurun_out/$d /tmp/$d; cd /tmp/$d
  echo "$R/tests/data/$f" | PDEODE_HOST_TIMING=1 PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver > console.txt 2> err.txt
  echo "$d exit $?"; grep "Time to compute" console.txt; tail -2 err.txt
  for x in total.dat COprod.dat peakInfo.dat statReport.dat console.txt; do [ -f $x ] && cp $x $R/gpurun_out/$d/; done
  [ -f output.dat ] && python $R/tools/frames_to_npz.py output.dat $R/gpurun_out/$d/frames.npz 10
done
