#!/bin/bash
# persistent-solve strip width on N GPUs: tools/persist_bs.sh N
N=${1:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for bs in 256 128 64; do
  echo "N=$N persist BS=$bs: $(PDEODE_PERSIST_BS=$bs timeout 300 $T --master-port 2956${bs:0:1} bench.py --gpus $N --steps 3 --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e cell-it/s  %.2f ms/step' % (d['value'], d['ms_per_step']))")"
done
