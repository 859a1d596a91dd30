#!/usr/bin/env python
"""Packs selected frames of an output.dat into a small .npz (M and C columns only).

    python tools/frames_to_npz.py output.dat out.npz [every]

Frames are separated by an empty line (what seperateOutput.py keys on); rows
inside a frame by a blank-only line.  `every` keeps frames 0, every, 2*every, ...
"""
import sys

import numpy as np


def main():
    path, out = sys.argv[1], sys.argv[2]
    every = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    frames = {}
    k = 0
    M, C = [], []
    keep = True
    with open(path) as f:
        for line in f:
            if line == "\n":                      # frame terminator (P:846)
                if keep:
                    frames[f"M{k}"] = np.array(M)
                    frames[f"C{k}"] = np.array(C)
                k += 1
                M, C = [], []
                keep = (k % every == 0)
                continue
            if not keep:
                continue
            t = line.split()
            if len(t) == 4:
                M.append(float(t[2]))
                C.append(float(t[3]))
    np.savez_compressed(out, nframes=k, **frames)
    print("frames", k, "kept", sorted(int(n[1:]) for n in frames if n[0] == "M"))


if __name__ == "__main__":
    main()
