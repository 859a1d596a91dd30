#!/bin/bash
# Round-2 record on ONE B200 (via gpurun): GPU parity tests + smoke, the bench line the driver
# will run, the ncu launch list of the same program, the 1024^2 full-run trace (production
# summation order; compared offline with the CPU oracle's trace by tools/fullrun_compare.py)
# and a one-GPU ensemble line.  Everything lands in gpurun_out/r02/.
# Usage: tools/r02_record.sh [tests] [bench] [ncu] [trace] [ensemble]   (default: all five)
WHAT=${*:-tests bench ncu trace ensemble}
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
for w in $WHAT; do case $w in
tests)
  timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -n 6 $O/pytest_gpu.log
  timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1
  echo "smoke exit $?"; tail -n 4 $O/smoke.log ;;
bench)
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_1gpu_4096.json 2> $O/bench_1gpu_4096.err
  echo "bench exit $?"; tail -n 12 $O/bench_1gpu_4096.err; cut -c1-600 $O/bench_1gpu_4096.json ;;
ncu)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-strong16384 --kernel-reps 5"
  timeout 300 $CMD > $O/ncu_plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 \
      --csv --log-file $O/launches_bench_4096.csv $CMD > $O/ncu_run.log 2>&1
  echo "ncu exit $?"; tail -n 2 $O/ncu_run.log; wc -l $O/launches_bench_4096.csv ;;
trace)
  timeout 300 python tools/fullrun_trace.py --impl gpu --grid 1024 --out $O/full1024_gpu_tree.npz > $O/trace1024.log 2>&1
  echo "trace exit $?"; tail -n 1 $O/trace1024.log ;;
ensemble)
  timeout 300 python bench.py --ensemble 32 --gpus 1 > $O/bench_ensemble32_1gpu.json 2> $O/bench_ensemble32_1gpu.err
  echo "ensemble exit $?"; tail -n 3 $O/bench_ensemble32_1gpu.err; cut -c1-400 $O/bench_ensemble32_1gpu.json ;;
esac; done
