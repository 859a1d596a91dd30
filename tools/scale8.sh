#!/bin/bash
# 8-GPU session (gpurun --gpus 8): correctness, 4096^2 scaling point, 16384^2 strong-scaling point.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29511 tools/dist_check.py --grid 1024 --steps 2 > gpurun_out/dist8.log 2>&1
echo "dist_check N=8 exit $?: $(grep -E 'exchange|DIST' gpurun_out/dist8.log | tr '\n' ' ')"
timeout 300 $T --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench8_4096.json 2> gpurun_out/bench8_4096.err
echo "bench N=8 4096 exit $?"; cut -c1-260 gpurun_out/bench8_4096.json
timeout 600 $T --master-port 29513 bench.py --gpus 8 --grid 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench8_16384.json 2> gpurun_out/bench8_16384.err
echo "bench N=8 16384 exit $?"; cut -c1-260 gpurun_out/bench8_16384.json; tail -2 gpurun_out/bench8_16384.err
