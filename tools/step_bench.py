#!/usr/bin/env python
"""Time per time step of small grids, per engine (GPU box).

    python tools/step_bench.py [--grids 257,512,1024] [--steps 300]

Shipped parameter values, MinitialCond = 5 style blobs; every engine starts
from the same state, runs 50 untimed steps and then `--steps` timed ones
(wall clock around pdeode_step calls: a step ends with a host wait anyway).
Prints one JSON line per (grid, engine).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))
sys.path.insert(0, ROOT)
import pdeode_b200 as pb  # noqa: E402
from bench import ic_default_like  # noqa: E402

ENGINES = {
    "whole-step": {"PDEODE_WHOLE_STEP": "1"},
    "persistent": {"PDEODE_WHOLE_STEP": "0", "PDEODE_PERSIST": "1"},
    "kernels": {"PDEODE_WHOLE_STEP": "0", "PDEODE_PERSIST": "0"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="257,512,1024")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--engines", default="whole-step,persistent,kernels")
    ap.add_argument("--rows", default="")
    ap.add_argument("--poll", default="")
    ap.add_argument("--groups", default="")
    ap.add_argument("--esoln", type=float, default=1e-8,
                    help=">= 2 makes every step an empty one (P:706 false on entry): launch + report only")
    args = ap.parse_args()
    for gs in args.grids.split(","):
        row = col = int(gs)
        M0, C0 = ic_default_like(row, col)
        for eng in args.engines.split(","):
            ws = eng == "whole-step"
            combos = [(r, po, gr) for r in (args.rows.split(",") if args.rows and ws else [""])
                      for po in (args.poll.split(",") if args.poll and ws else [""])
                      for gr in (args.groups.split(",") if args.groups and ws else [""])]
            for rows, po, gr in combos:
                for k, flag in (("PDEODE_WHOLE_STEP", 1), ("PDEODE_PERSIST", 1), ("PDEODE_STEP_ROWS", rows),
                                ("PDEODE_STEP_POLL", po),
                                ("PDEODE_STEP_GROUPS", gr)):
                    if flag:                      # options not swept stay as the caller's environment has them
                        os.environ.pop(k, None)
                os.environ.update(ENGINES[eng])
                for key, val in (("PDEODE_STEP_ROWS", rows), ("PDEODE_STEP_POLL", po),
                                 ("PDEODE_STEP_GROUPS", gr)):
                    if val:
                        os.environ[key] = val
                s = pb.Solver(row, col, pb.default_params(row, eSoln=args.esoln))
                s.set_state(M0, C0)
                for _ in range(50):
                    s.step()
                t0 = time.perf_counter()
                its = pic = 0
                for _ in range(args.steps):
                    r = s.step()
                    its += r["nitSum"]
                    pic += r["countIters"]
                dt = time.perf_counter() - t0
                M, _ = s.get_state()
                print(json.dumps({"grid": row, "engine": eng, "rows": rows, "poll": po, "groups": gr, "eSoln": args.esoln,
                                  "us_per_step": round(1e6 * dt / args.steps, 1),
                                  "picard_per_step": round(pic / args.steps, 2),
                                  "cg_per_step": round(its / args.steps, 2),
                                  "sumM": float(M.sum()), "info": s.engine_info()}), flush=True)
                s.close()


if __name__ == "__main__":
    main()
