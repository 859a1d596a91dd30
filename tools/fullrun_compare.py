#!/usr/bin/env python
"""Compares two traces of tools/fullrun_trace.py (reference side first):
  * bit identity of M and C at every frame (sha256 of the arrays),
  * per-step Picard counts and CG-iteration sums: how many steps differ, by how much,
  * relative L2 distance of the (decimated) fields at every frame.
Usage: fullrun_compare.py REF.npz OTHER.npz [--label text]   (markdown on stdout)"""
import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ref")
    ap.add_argument("other")
    ap.add_argument("--label", default="")
    ap.add_argument("--steps", type=int, default=0,
                    help="compare the first N steps only (a run still under way: its <out>.ckpt.npz)")
    a = ap.parse_args()
    r = np.load(a.ref, allow_pickle=True)
    o = np.load(a.other, allow_pickle=True)
    n = min(len(r["count"]), len(o["count"]))
    if "k" in r.files:                       # a checkpoint: steps done so far
        n = min(n, int(r["k"]))
    if "k" in o.files:
        n = min(n, int(o["k"]))
    if a.steps > 0:
        n = min(n, a.steps)
    nfr = min(len(r["frames_at"]), len(o["frames_at"]),
              int((np.asarray(r["frames_at"]) <= n).sum()), int((np.asarray(o["frames_at"]) <= n).sum()))
    grid = int(r["grid"])
    print(f"### {a.label or (a.other + ' vs ' + a.ref)}")
    print(f"grid {grid} x {grid}, {n} time steps, {nfr} frames; "
          f"reference side: {str(r['info'])}; other side: {str(o['info'])}\n")
    same_frames = [(str(x) == str(y)) and (str(u) == str(v))
                   for x, y, u, v in zip(r["hashM"][:nfr], o["hashM"][:nfr], r["hashC"][:nfr], o["hashC"][:nfr])]
    print(f"* frames bit-identical (sha256 of M and of C): **{sum(same_frames)} of {len(same_frames)}**"
          + ("" if all(same_frames) else f" (first difference at frame {same_frames.index(False)}, "
                                          f"step {int(r['frames_at'][same_frames.index(False)])})"))
    dc = o["count"][:n].astype(np.int64) - r["count"][:n]
    dn = o["nitsum"][:n].astype(np.int64) - r["nitsum"][:n]
    print(f"* Picard count per step: identical in {int((dc == 0).sum())} of {n} steps"
          + ("" if (dc == 0).all() else f" (max |difference| {int(np.abs(dc).max())})"))
    solves = np.maximum(r["count"][:n].astype(np.int64), 1)
    out2 = np.abs(dn) > 2 * solves
    print(f"* CG iterations per step: identical in {int((dn == 0).sum())} of {n} steps "
          f"({100.0 * (dn == 0).mean():.2f} %); |difference| <= 2 per solve in {n - int(out2.sum())} "
          f"({100.0 * (1 - out2.mean()):.3f} %); outside that: {int(out2.sum())} steps "
          f"({100.0 * out2.mean():.3f} %), largest |difference| {int(np.abs(dn).max())} in one step")
    hist = {k: int((np.abs(dn) == k).sum()) for k in range(0, 9)}
    hist[">8"] = int((np.abs(dn) > 8).sum())
    print(f"* histogram of |CG difference| per step: {hist}")
    print(f"* totals: Picard {int(r['count'][:n].sum())} vs {int(o['count'][:n].sum())}, "
          f"CG {int(r['nitsum'][:n].sum())} vs {int(o['nitsum'][:n].sum())}, "
          f"max CG per solve {int(r['nitmax'][:n].max())} vs {int(o['nitmax'][:n].max())}")
    first = int(np.argmax(dn != 0)) if (dn != 0).any() else -1
    print(f"* first step whose CG count differs: {first if first >= 0 else 'none'}")
    eM = [float(np.linalg.norm(x - y) / max(np.linalg.norm(x), 1e-300)) for x, y in zip(r["M"][:nfr], o["M"][:nfr])]
    eC = [float(np.linalg.norm(x - y) / max(np.linalg.norm(x), 1e-300)) for x, y in zip(r["C"][:nfr], o["C"][:nfr])]
    print(f"* relative L2 of the fields over the frames (every {int(r['stride'])}-th point): "
          f"M max {max(eM):.2e}, C max {max(eC):.2e}\n")
    print("| frame | step | relL2 M | relL2 C | bit-identical |")
    print("|---:|---:|---:|---:|:---:|")
    nf = len(eM)
    for k in sorted(set(list(range(0, nf, 10)) + [nf - 1])):
        print(f"| {k} | {int(r['frames_at'][k])} | {eM[k]:.2e} | {eC[k]:.2e} | {'yes' if same_frames[k] else 'no'} |")
    print()


if __name__ == "__main__":
    main()
