#!/bin/bash
# Round-2 record, part 2 (one B200): per-kernel times at 4096^2 with the library's own launch
# shapes, time per step of the small grids per engine, the shipped 257^2 run end to end
# through the drop-in program, and one `ncu --set full` capture of the two CG kernels.
R=${GRAFT_REPO_ROOT:-$PWD}
O=$R/gpurun_out/r02
mkdir -p $O
timeout 100 python tools/sweep_kernels.py --reps 40 \
  --kernels cg_spmv_x,cg_update_r,cg_spmv,cg_update,init,solvec,diffcoef > $O/kernel_times_4096.jsonl 2> $O/kernel_times.err
echo "kernels exit $?"; cat $O/kernel_times_4096.jsonl
timeout 100 python tools/step_bench.py --grids 257,512,1024 --steps 300 > $O/step_bench.jsonl 2> $O/step_bench.err
echo "step_bench exit $?"; cat $O/step_bench.jsonl
mkdir -p /tmp/run257 && cd /tmp/run257
( time ( echo "$R/tests/data/shipped257.txt" | PDEODE_HOST_TIMING=1 PDEODE_SEED=1 timeout 100 $R/pde-odesolver_b200/bin/pdeode_solver > console.txt 2> err.txt ) ) 2> time.txt
echo "run257 exit $?"; tail -4 console.txt; tail -3 err.txt; cat time.txt
mkdir -p $O/run257; cp console.txt err.txt time.txt statReport.dat $O/run257/ 2>/dev/null; ls -la > $O/run257/files.txt
cd $R
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-strong16384 --kernel-reps 5"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_bulk|k_update_r' -s 200 -c 2 \
    -o $O/ncu_full_cg -f $CMD > $O/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -n 2 $O/ncu_full.log | cut -c1-300; ls -la $O/*.ncu-rep
