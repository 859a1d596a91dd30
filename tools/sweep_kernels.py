#!/usr/bin/env python
"""Times the path's kernels in isolation over launch configurations (GPU box).

    python tools/sweep_kernels.py [--grid 4096] [--reps 30] [--bs 64,128,256] [--tr 8,16,32,64]

Each configuration gets its own context (the launch shape is read from
PDEODE_BS / PDEODE_TR / PDEODE_VARIANT at pdeode_create).  Prints one line per
configuration: kernel, ms per launch, algorithmic GB/s, fraction of the
measured HBM peak.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))
sys.path.insert(0, ROOT)
import pdeode_b200 as pb  # noqa: E402
from bench import ic1, measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--bs", default="0")
    ap.add_argument("--tr", default="0")
    ap.add_argument("--variant", default="")
    ap.add_argument("--kernels", default="cg_spmv,cg_update")
    args = ap.parse_args()
    peak, _ = measured_peak()
    row = col = args.grid
    n = row * col
    M0, C0 = ic1(row, col, 0.9, 0.3)
    kid = {"cg_spmv": pb.KERNEL_CG_SPMV, "cg_update": pb.KERNEL_CG_UPDATE,
           "cg_spmv_x": pb.KERNEL_CG_SPMV_X, "cg_update_r": pb.KERNEL_CG_UPDATE_R,
           "init": pb.KERNEL_INIT, "solvec": pb.KERNEL_SOLVEC, "diffcoef": pb.KERNEL_DIFFCOEF}
    variants = args.variant.split(",") if args.variant else [""]
    for var in variants:
        for bs in args.bs.split(","):
            for tr in args.tr.split(","):
                # "0" = leave the library's own heuristic in charge
                for key, val in (("PDEODE_BS", bs), ("PDEODE_TR", tr), ("PDEODE_VARIANT", var)):
                    if val and val != "0":
                        os.environ[key] = val
                    else:
                        os.environ.pop(key, None)
                s = pb.Solver(row, col, pb.default_params(row, eSoln=1.5))
                s.set_state(M0, C0)
                s.set_maxit(3)
                s.step()                       # realistic d, ac, r, p, q
                for k in args.kernels.split(","):
                    ms = s.time_kernel(kid[k], args.reps)
                    gbs = pb.Solver.bytes_per_cell(kid[k]) * n / ms / 1e6
                    print(json.dumps({"variant": var, "bs": int(bs), "tr": int(tr), "kernel": k,
                                      "ms": round(ms, 5), "GBps": round(gbs, 1),
                                      "frac": round(gbs / peak, 4)}), flush=True)
                s.close()


if __name__ == "__main__":
    main()
