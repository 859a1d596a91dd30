#!/bin/bash
# Phase times of the whole-step launch as CTA 0 sees them (ns | SM clocks per phase):
# D, centre diagonal + residual, then per CG iteration SpMV phase and update phase, then
# C update + next diagonal; each ends in one grid-wide sum.
for g in 257 512 1024; do
  echo "grid $g"
  PDEODE_STEP_STAMPS=1 timeout 120 python tools/step_bench.py --grids $g --steps 3 --engines whole-step 2>&1 | grep "step stamps" | tail -2
done
