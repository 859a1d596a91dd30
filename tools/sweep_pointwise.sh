#!/bin/bash
# Launch-shape sweep of the pointwise kernels of a time step at 4096^2 (GPU box):
# D(M) (k_diffcoef), the C update (k_solvec), centre diagonal + first residual (k_init_bulk).
# One process per setting: the knobs are read once per process.
out=${1:-gpurun_out/r02_sweep_pointwise.jsonl}
: > $out
for u in 1 2 4; do for c in 8 16; do
  echo "# diffcoef unroll=$u stream_ctas=$c" >> $out
  PDEODE_DIFFCOEF_UNROLL=$u PDEODE_STREAM_CTAS=$c python tools/sweep_kernels.py --kernels diffcoef --reps 40 >> $out 2>/dev/null
done; done
for u in 1 2; do for c in 4 8 16; do
  echo "# solvec unroll=$u stream_ctas=$c" >> $out
  PDEODE_SOLVEC_UNROLL=$u PDEODE_STREAM_CTAS=$c python tools/sweep_kernels.py --kernels solvec --reps 40 >> $out 2>/dev/null
done; done
for d in 2 4; do for bs in 64 128 256; do for tr in 8 16 32 64; do
  echo "# init depth=$d bs=$bs tr=$tr" >> $out
  PDEODE_INIT_DEPTH=$d PDEODE_INIT_BS=$bs PDEODE_INIT_TR=$tr python tools/sweep_kernels.py --kernels init --reps 40 >> $out 2>/dev/null
done; done; done
cat $out
