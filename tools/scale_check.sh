#!/bin/bash
# Multi-GPU check + strong-scaling bench on N GPUs (GPU box, via gpurun --gpus N).
# Usage: tools/scale_check.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for P in 0 1; do
  PDEODE_PERSIST=$P timeout 300 $T --master-port 2951$P tools/dist_check.py --grid 1024 --steps 2 > gpurun_out/dist${N}_p$P.log 2>&1
  echo "dist_check N=$N persist=$P exit $?: $(grep -E 'exchange|DIST' gpurun_out/dist${N}_p$P.log | tr '\n' ' ')"
done
for P in 0 1 auto; do
  if [ $P = auto ]; then unset PDEODE_PERSIST; else export PDEODE_PERSIST=$P; fi
  timeout 400 $T --master-port 2952${#P} bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench${N}_p$P.json 2> gpurun_out/bench${N}_p$P.err
  echo "bench N=$N persist=$P exit $?: $(python -c "import sys,json; d=json.loads(open('gpurun_out/bench${N}_p$P.json').read()); print('%.3e cell-it/s  %.2f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value'] if d['e2e'] else 0))" 2>&1 | tail -1)"
done
