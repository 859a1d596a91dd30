#!/usr/bin/env python
"""Time per CG iteration of the persistent solve on one slab-shaped grid (one GPU).

    python tools/slab_bench.py [--rows 512] [--cols 4096] [--bs 64,128,256]

A rank of an N-GPU run of a 4096^2 grid owns 4096/N rows; this runs a grid of
that shape alone, so tile shapes and pipeline fill / drain are those of the
multi-GPU run and only the NVLink exchange is missing.  Prints us per CG
iteration and the streaming time the same cells take at the measured peak.
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
from bench import ic1, measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=512)
    ap.add_argument("--cols", type=int, default=4096)
    ap.add_argument("--bs", default="0,64,128,256")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--ngpus", type=int, default=1,
                    help="row slabs over this many GPUs of one process (pdeode_create_group): "
                         "--rows is then the whole grid, e.g. --ngpus 2 --rows 1024 gives the two "
                         "ranks the 512-row slabs of an 8-GPU run of 4096^2, exchange included")
    ap.add_argument("--prefetch", default="",
                    help="comma list of PDEODE_PERSIST_PREFETCH settings to compare (e.g. 0,1)")
    ap.add_argument("--persist-only", action="store_true")
    args = ap.parse_args()
    peak, _ = measured_peak()
    row, col = args.rows, args.cols
    M0, C0 = ic1(row, col, 0.9, 0.3)
    os.environ["PDEODE_WHOLE_STEP"] = "0"
    for persist in (("1",) if args.persist_only else ("1", "0")):
      for pf in (args.prefetch.split(",") if (args.prefetch and persist == "1") else [""]):
        for bs in (args.bs.split(",") if persist == "1" else ["0"]):
            os.environ["PDEODE_PERSIST"] = persist
            os.environ.pop("PDEODE_PERSIST_BS", None)
            os.environ.pop("PDEODE_PERSIST_PREFETCH", None)
            if pf not in ("", "auto"):
                os.environ["PDEODE_PERSIST_PREFETCH"] = pf
            if bs != "0":
                os.environ["PDEODE_PERSIST_BS"] = bs
            s = pb.Solver(row, col, pb.default_params(col), ngpus=args.ngpus)  # xDel of the 4096^2 problem
            s.set_state(M0, C0)
            s.set_maxit(400)
            s.step()
            t0 = time.perf_counter()
            its = sum(s.step()["nitSum"] for _ in range(args.steps))
            dt = time.perf_counter() - t0
            print(json.dumps({"rows": row, "cols": col, "persist": persist, "persist_bs": bs,
                              "us_per_cg_iteration": round(1e6 * dt / its, 2), "cg_iterations": its,
                              "streaming_us_at_peak": round(96.0 * row * col / args.ngpus / peak / 1e3, 2),
                              "ngpus": args.ngpus,
                              "info": s.engine_info()}), flush=True)
            s.close()


if __name__ == "__main__":
    main()
