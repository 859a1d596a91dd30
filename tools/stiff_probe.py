#!/usr/bin/env python
"""The stiff sub-record of bench.py (beta = 6, height 0.95, one Picard iterate, CG cut at a cap)
step by step at a chosen grid, with wall times -- to see what a step of that leg does."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))
import pdeode_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--cap", type=int, default=600)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--beta", type=int, default=6)
    ap.add_argument("--height", type=float, default=0.95)
    ap.add_argument("--esoln", type=float, default=1.5)
    args = ap.parse_args()
    g = args.grid
    s = pb.Solver(g, g, pb.default_params(g, beta=args.beta, eSoln=args.esoln))
    s.set_maxit(args.cap)
    print(s.engine_info(), flush=True)
    for k in range(args.steps):
        t0 = time.perf_counter()
        s.set_initial_condition(1, 0.3, args.height)
        t1 = time.perf_counter()
        r = s.step_trace()
        t2 = time.perf_counter()
        print(f"step {k}: IC {t1 - t0:.3f} s, step {t2 - t1:.3f} s, rc {r['rc']} Picard {r['countIters']} "
              f"nits {r['nits'][:8]} diffs {r['diffs'][:4]} mass {s.mass()}", flush=True)
    s.close()


if __name__ == "__main__":
    main()
