#!/usr/bin/env python
"""Compares two runs of the host program (e.g. GPU vs CPU-oracle-driven).

    python tools/compare_runs.py GPU_DIR CPU_DIR

Each directory holds total.dat and frames.npz (tools/frames_to_npz.py).  Prints,
per kept frame, the relative L2 difference of M and C, and the largest absolute
difference of the mean-M / mean-C columns of total.dat over the samples both
runs have.
"""
import sys

import numpy as np


def main():
    g, c = sys.argv[1], sys.argv[2]
    fg, fc = np.load(f"{g}/frames.npz"), np.load(f"{c}/frames.npz")
    ks = sorted(int(n[1:]) for n in fg.files if n[0] == "M" and n in fc.files)
    worst = 0.0
    for k in ks:
        out = []
        for name in ("M", "C"):
            a, b = fg[f"{name}{k}"], fc[f"{name}{k}"]
            if a.shape != b.shape:
                out.append(f"{name}: shapes {a.shape} vs {b.shape}")
                continue
            nb = np.linalg.norm(b)
            e = np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
            worst = max(worst, e)
            out.append(f"relL2({name}) = {e:.3e}")
        print(f"frame {k:3d}: " + "  ".join(out))
    def table(path):                     # a run still in progress ends in a partial line
        rows = [l.split() for l in open(path)]
        return np.array([[float(v) for v in r] for r in rows if len(r) == 3])
    tg, tc = table(f"{g}/total.dat"), table(f"{c}/total.dat")
    m = min(len(tg), len(tc))
    print(f"total.dat: {m} common samples; max |d mean M| = {np.abs(tg[:m, 1] - tc[:m, 1]).max():.3e}, "
          f"max |d mean C| = {np.abs(tg[:m, 2] - tc[:m, 2]).max():.3e}; final mean M {tg[m - 1, 1]:.6f}")
    print(f"worst frame difference: {worst:.3e}")


if __name__ == "__main__":
    main()
