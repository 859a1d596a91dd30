#!/usr/bin/env python
"""Full run of the shipped configuration (parameter.txt: 30 001 steps, frames every 333),
stepped directly through the C ABI (GPU) or the serial CPU oracle, recording
  * per time step: Picard count, CG iteration sum and maximum          (P:722-723, P:739-740)
  * per frame (state BEFORE the step, P:683-687): sha256 of the M and C arrays (bit identity),
    their means, and the fields decimated by --stride (for relative-L2 comparisons offline)
into one compressed .npz.  compare: tools/fullrun_compare.py.

  python tools/fullrun_trace.py --impl oracle --grid 257 --out gpurun_out/full257_oracle.npz
  PDEODE_SUM_ORDER=serial python tools/fullrun_trace.py --impl gpu --grid 257 --out ...

Test infrastructure (it can drive the oracle)."""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["oracle", "gpu"], required=True)
    ap.add_argument("--grid", type=int, default=257)
    ap.add_argument("--steps", type=int, default=30001)       # counter = 0 .. 30000 (P:663)
    ap.add_argument("--frame-every", type=int, default=333)   # filter = int(tEnd/(nOuts*tDel)) (P:638)
    ap.add_argument("--stride", type=int, default=0)          # 0: 2 for <= 512, else 8
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--perturb", type=float, default=0.0,
                    help="start from M0 * (1 +- eps), a pseudo-random sign per cell: the size of a "
                         "rounding difference; shows how far the SAME code drifts from such a start")
    ap.add_argument("--omp", type=int, default=0,
                    help="oracle only: the OpenMP build (the reference's runAll_paral.sh variant) on "
                         "this many threads instead of the serial build")
    ap.add_argument("--out", required=True)
    ap.add_argument("--checkpoint", action="store_true",
                    help="write <out>.ckpt.npz (trace so far + full state) at every frame and "
                         "continue from it when it exists: (M, C) at a step boundary is the whole "
                         "state of solveOrder2 (P:703-704 copy Mnew, Cnew over M, C)")
    args = ap.parse_args()
    import oracle_lib as ol
    row = col = args.grid
    stride = args.stride or (2 if row <= 512 else 8)
    M0, C0 = ol.ic_blobs(row, col, 0.02, 0.0390625, 60, args.seed)   # MinitialCond = 5, fixed seed (D5)
    if args.perturb != 0.0:
        sign = np.where(np.random.RandomState(12345).randint(0, 2, M0.size) == 1, 1.0, -1.0)
        M0 = M0 * (1.0 + args.perturb * sign)
    if args.impl == "oracle":
        ol.build_oracle()
        if args.omp > 0:
            os.environ["OMP_NUM_THREADS"] = str(args.omp)
            s = ol.Oracle(row, col, ol.default_params(row), omp=True)
            info = f"OpenMP CPU oracle on {args.omp} threads (the reference's runAll_paral.sh build, restated)"
        else:
            s = ol.Oracle(row, col, ol.default_params(row))
            info = "serial CPU oracle (bit-identical to the reference's translated source)"
    else:
        import pdeode_b200 as pb
        s = pb.Solver(row, col, pb.default_params(row))
        info = " ".join(f"{k}={v}" for k, v in s.engine_info().items())
    s.set_state(M0, C0)
    count = np.zeros(args.steps, np.int32)
    nitsum = np.zeros(args.steps, np.int32)
    nitmax = np.zeros(args.steps, np.int32)
    frames_at, hM, hC, mM, mC, fM, fC = [], [], [], [], [], [], []
    t0 = time.time()
    k0 = 0
    ckpt = args.out + ".ckpt.npz"
    if args.checkpoint and os.path.exists(ckpt):
        z = np.load(ckpt)
        k0 = int(z["k"])
        count[:k0] = z["count"][:k0]; nitsum[:k0] = z["nitsum"][:k0]; nitmax[:k0] = z["nitmax"][:k0]
        frames_at = list(z["frames_at"]); hM = list(z["hashM"]); hC = list(z["hashC"])
        mM = list(z["meanM"]); mC = list(z["meanC"]); fM = list(z["M"]); fC = list(z["C"])
        s.set_state(z["stateM"], z["stateC"])
        s.set_mnew(z["stateM"])          # warm start of the next solve (P:714): Mnew == M here
        t0 -= float(z["wall_s"])
        print(f"continuing from step {k0}", file=sys.stderr, flush=True)
    for k in range(k0, args.steps):
        if k % args.frame_every == 0:
            M, C = s.get_state()
            if args.checkpoint and k > k0:
                tmp = ckpt + ".tmp.npz"
                np.savez_compressed(tmp, k=k, grid=row, stride=stride, impl=args.impl, info=info,
                                    count=count, nitsum=nitsum, nitmax=nitmax,
                                    frames_at=np.array(frames_at), hashM=np.array(hM), hashC=np.array(hC),
                                    meanM=np.array(mM), meanC=np.array(mC), M=np.array(fM), C=np.array(fC),
                                    stateM=M, stateC=C, wall_s=time.time() - t0)
                os.replace(tmp, ckpt)
            frames_at.append(k)
            hM.append(hashlib.sha256(M.tobytes()).hexdigest())
            hC.append(hashlib.sha256(C.tobytes()).hexdigest())
            mM.append(M.mean()); mC.append(C.mean())
            fM.append(M.reshape(row, col)[::stride, ::stride].copy())
            fC.append(C.reshape(row, col)[::stride, ::stride].copy())
            print(f"step {k}: {time.time() - t0:.0f} s, CG so far {int(nitsum[:k].sum())}", file=sys.stderr, flush=True)
        r = s.step()
        count[k] = r["countIters"]; nitsum[k] = r["nitSum"]; nitmax[k] = r["nitMax"]
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    np.savez_compressed(args.out, grid=row, stride=stride, impl=args.impl, info=info, count=count,
                        nitsum=nitsum, nitmax=nitmax, frames_at=np.array(frames_at),
                        hashM=np.array(hM), hashC=np.array(hC), meanM=np.array(mM), meanC=np.array(mC),
                        M=np.array(fM), C=np.array(fC), wall_s=time.time() - t0)
    print(f"{args.impl} {row}^2: {args.steps} steps in {time.time() - t0:.1f} s, "
          f"Picard avg {count.mean():.5f}, CG per solve {nitsum.sum() / count.sum():.4f}, max {nitmax.max()}")


if __name__ == "__main__":
    main()
