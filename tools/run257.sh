#!/bin/bash
# The reference's shipped configuration end to end on the GPU box:
# 257^2 (runAll.sh mode) and 257x4 (twoDimension_runall.sh mode), 30001 steps each.
R=$GRAFT_REPO_ROOT
for tag in run257:shipped257.txt run257w:shipped257_wave.txt; do
  d=${tag%%:*}; f=${tag##*:}
  mkdir -p $R/gpurun_out/$d /tmp/$d; cd /tmp/$d
  s=$(date +%s.%N)
  echo "$R/tests/data/$f" | PDEODE_SEED=1 $R/pde-odesolver_b200/bin/pdeode_solver > console.txt 2> err.txt
  e=$(date +%s.%N)
  echo "$d exit $?"; tail -11 console.txt | head -12; tail -2 err.txt
  ls -la | head -12
  for x in total.dat COprod.dat peakInfo.dat statReport.dat console.txt 2D_output.dat; do [ -f $x ] && cp $x $R/gpurun_out/$d/; done
done
