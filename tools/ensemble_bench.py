#!/usr/bin/env python
"""Parameter-sweep ensemble (BASELINE config 5) on the GPUs of this box.

    python tools/ensemble_bench.py [--members 32] [--grid 512] [--steps 20] [--per-gpu 1,4,8]

Writes numbered parameter files 1.txt .. N.txt (the reference's positional
format; delta, nu, kappa, gama, alpha, beta and height vary per member), runs
pde-odesolver_b200/bin/pdeode_ensemble on them and reports the aggregate
grid-cell CG-iterations/s (from each member's statReport.dat) per setting.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TEMPLATE = """! ensemble member {k}
! positional parameter file (names are ignored by the reader)
! ---
pSize     = {grid}
true2D    = 0
depth     = 0.3
height    = {height}
innocPts  = 0
alpha     = {alpha}
beta      = {beta}
nu        = {nu}
kappa     = {kappa}
gama      = {gama}
delta     = {delta}
nOuts     = 1
tEnd      = {tEnd}
tDel      = 0.001
xDel      = 1/real(row)
e1        = 1.0e-8
e2        = 1.0e-8
eSoln     = 1.0e-8
nit       = n*100
! ---
! ---
! ---
fSelect   = 1
dSelect   = 3
gSelect   = 1
MinitialCond = 1
"""


def member_params(k, grid, steps):
    # a deterministic spread around the shipped values
    return dict(k=k, grid=grid, height=0.70 + 0.02 * (k % 10), alpha=4, beta=3 + (k % 2),
                nu=0.25 + 0.01 * (k % 7), kappa=0.008 + 0.001 * (k % 5), gama=0.5 + 0.02 * (k % 6),
                delta=1e-4 * (1 + 0.1 * (k % 4)), tEnd=steps * 0.001 - 0.0005)


def stat_value(text, label):
    line = [l for l in text.splitlines() if label in l][0]
    return float(line.split("=")[1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=32)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--per-gpu", default="1,4,8")
    ap.add_argument("--persist", default="auto,0")
    args = ap.parse_args()
    exe = os.path.join(ROOT, "pde-odesolver_b200", "bin", "pdeode_ensemble")
    n = args.grid * args.grid
    for persist in args.persist.split(","):
        for per_gpu in args.per_gpu.split(","):
            with tempfile.TemporaryDirectory() as d:
                files = []
                for k in range(1, args.members + 1):
                    f = f"{k}.txt"
                    with open(os.path.join(d, f), "w") as fh:
                        fh.write(TEMPLATE.format(**member_params(k, args.grid, args.steps)))
                    files.append(f)
                env = dict(os.environ, PDEODE_MEMBERS_PER_GPU=per_gpu, PDEODE_SEED="1")
                env.pop("PDEODE_PERSIST", None)
                if persist != "auto":
                    env["PDEODE_PERSIST"] = persist
                t0 = time.perf_counter()
                r = subprocess.run([exe] + files, cwd=d, env=env, capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    print(json.dumps({"error": r.stdout[-400:] + r.stderr[-400:]}))
                    continue
                cg = 0.0
                steps = 0
                for k in range(1, args.members + 1):
                    txt = open(os.path.join(d, f"{k}Output", "statReport.dat")).read()
                    avg_iters = stat_value(txt, "Avg Iters for iterating")
                    avg_nit = stat_value(txt, "Avg Iters for linear")
                    con = open(os.path.join(d, f"{k}Output", "console.txt")).read()
                    nsteps = args.steps
                    cg += avg_nit * avg_iters * nsteps
                    steps += nsteps
                print(json.dumps({"members": args.members, "grid": args.grid, "steps_per_member": args.steps,
                                  "members_per_gpu": int(per_gpu), "persist": persist,
                                  "wall_s": round(dt, 3), "cg_iterations": round(cg),
                                  "cell_iterations_per_s": n * cg / dt}), flush=True)


if __name__ == "__main__":
    main()
