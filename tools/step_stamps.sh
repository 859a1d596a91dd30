#!/bin/bash
# phase times of the whole-step launch (CTA 0) for a few steps
mkdir -p gpurun_out
for gr in 1 3; do
echo "groups=$gr"
PDEODE_STEP_GROUPS=$gr PDEODE_STEP_STAMPS=1 timeout 120 python tools/step_bench.py --grids 257 --steps 3 --engines whole-step 2>&1 | grep -A1 "step stamps" | tail -4
done
