#!/usr/bin/env python
"""Golden vectors produced BY THE REFERENCE (not by the oracle): oracle/_ref/libref.so is the
reference's own source, /root/reference/*.f90, translated statement by statement to C
(oracle/f90c/f90toc.py; this image has no Fortran compiler) and compiled with the flags of
runAll.sh:6.  Each case calls the reference's solveOrder2 (PDE-ODEsolver.f90:600-763) with
tEnd = (s - 0.5)*tDel for s = 2 .. K, i.e. runs s time steps (a one-step run is impossible: the
reference's frame filter int(tEnd/(nOuts*tDel)) would be 0 and MOD(counter, 0) traps);
the statistics it returns (P:722-723, P:739-745) give the cumulative Picard count and CG-iteration
sum after s steps -- hence those of every single step from the third on -- and its M, C arguments
the state after s steps.

    python tests/golden/make_ref_golden.py        # rewrites tests/golden/ref_*.npz

The files are small and committed: they travel to the GPU box (where /root/reference does not
exist) and pin both the oracle (tests/test_golden_ref.py, CPU) and the CUDA path
(tests/test_gpu_parity.py, GPU) to the reference's own arithmetic.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol   # noqa: E402  (initial conditions and the parameter struct only)
import ref_lib as rl      # noqa: E402

# name, row, col, true2D, parameter overrides, initial condition, K steps
CASES = [
    ("stiff33", 33, 33, 0, dict(), ("ic1", 0.9, 0.3), 12),
    ("stiff65", 65, 65, 0, dict(), ("ic1", 0.9, 0.3), 6),
    ("wave48x4", 48, 4, 1, dict(), ("ic1", 0.9, 0.3), 20),
    ("wave257x4", 257, 4, 1, dict(), ("ic1", 0.9, 0.3), 10),
    ("blobs33", 33, 33, 0, dict(tDel=0.05), ("blobs", 0.3, 0.2, 6, 3), 16),
    ("rect40x24_f6d2", 40, 24, 0, dict(fSelect=6, dSelect=2, tDel=0.01), ("ic1", 0.8, 0.4), 12),
    ("beta6_25", 25, 25, 0, dict(beta=6), ("ic1", 0.93, 0.3), 4),
    ("shipped65", 65, 65, 0, dict(), ("blobs", 0.02, 0.0390625, 60, 7), 8),
]

PKEYS = ("delta", "nu", "kappa", "gama", "tDel", "xDel", "e1", "e2", "eSoln", "alpha", "beta",
         "dSelect", "fSelect", "gSelect")


def state(row, col, ic):
    if ic[0] == "ic1":
        return ol.ic1(row, col, ic[1], ic[2])
    return ol.ic_blobs(row, col, ic[1], ic[2], ic[3], ic[4])


def main():
    ref = rl.Ref()
    for name, row, col, true2D, over, ic, K in CASES:
        p = ol.default_params(row, **over)
        M0, C0 = state(row, col, ic)
        Ms, Cs, totP, totN, maxP, maxN = [], [], [], [], [], []
        nsteps = list(range(2, K + 1))
        for steps in nsteps:
            with tempfile.TemporaryDirectory() as tmp:
                # counter*tDel <= tEnd holds for counter = 0 .. steps-1 (P:663); filter = steps-1 >= 1
                r = ref.solve_order2(M0, C0, row, col, p, (steps - 0.5) * p.tDel, 1, true2D, tmp)
            Ms.append(r["M"]); Cs.append(r["C"])
            totP.append(int(round(r["avgIters"] * steps)))               # P:745
            totN.append(int(round(r["avgNit"] * totP[-1])))              # P:744
            maxP.append(int(r["maxIters"])); maxN.append(int(r["maxNit"]))
        count = np.array(totP, np.int64)          # cumulative after nsteps[k] steps
        nitsum = np.array(totN, np.int64)
        out = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(
            out, row=row, col=col, true2D=true2D, params=np.array([getattr(p, k) for k in PKEYS], np.float64),
            M0=M0, C0=C0, nsteps=np.array(nsteps, np.int64), M=np.array(Ms), C=np.array(Cs), count=count, nitsum=nitsum,
            maxcount=np.array(maxP, np.int64), maxnit=np.array(maxN, np.int64))
        print(f"{name}: {row}x{col}, {K} steps, cumulative Picard {list(count)}, CG {list(nitsum)}, "
              f"{os.path.getsize(out) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
