#!/bin/bash
# quick multi-GPU check: both engines against the 1-GPU fields, then the 4096^2 bench
# usage: tools/scale_quick.sh N [nobench]
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for P in 1 0; do
  PDEODE_PERSIST=$P timeout 200 $T --master-port 2951$P tools/dist_check.py --grid 1024 --steps 2 > gpurun_out/dist${N}_p$P.log 2>&1
  echo "dist_check N=$N persist=$P exit $?: $(grep -E 'exchange|DIST' gpurun_out/dist${N}_p$P.log | tr '\n' ' ')"
  [ "$2" = nobench ] && [ $P = 1 ] && break
done
[ "$2" = nobench ] && exit 0
timeout 300 $T --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/bench${N}_4096.json 2> gpurun_out/bench${N}_4096.err
echo "bench N=$N 4096 exit $?"; cut -c1-200 gpurun_out/bench${N}_4096.json; tail -2 gpurun_out/bench${N}_4096.err
