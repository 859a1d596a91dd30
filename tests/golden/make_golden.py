#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/liboracle.so).

The reference ships no golden vectors and cannot be built in this image (no
Fortran compiler), so these fixtures freeze the oracle's own output: they
catch drift in the oracle and give the GPU tests a fixed target that does not
depend on building anything.  Re-run only when the oracle changes on purpose:

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

CASES = {
    # name: (row, col, param overrides, initial condition, steps)
    "blobs33": (33, 33, {}, ("blobs", 0.25, 0.2, 6, 5), 5),
    "wave48x4": (48, 4, {}, ("ic1", 0.9, 0.3), 5),            # true2D layout (col = 4)
    "rect40x24_f6d2": (40, 24, dict(fSelect=6, dSelect=2, alpha=2), ("ic1p", 0.6, 0.5), 3),
    "stiff33": (33, 33, dict(beta=5), ("ic1", 0.9, 0.3), 3),   # ~940 CG iterations in the first solve
}


def initial(kind, row, col):
    if kind[0] == "blobs":
        return ol.ic_blobs(row, col, kind[1], kind[2], kind[3], kind[4])
    if kind[0] == "ic1":
        return ol.ic1(row, col, kind[1], kind[2])
    if kind[0] == "ic1p":
        M, C = ol.ic1(row, col, kind[1], kind[2])
        rng = np.random.RandomState(9)
        M = M * (1.0 + 0.05 * rng.rand(M.size))
        C = 0.5 + 0.5 * rng.rand(M.size)
        return M, C
    raise ValueError(kind)


NEAR, NEAR_CAP, replay_step = ol.NEAR, ol.NEAR_CAP, ol.replay_step


def run(name):
    row, col, over, kind, steps = CASES[name]
    p = ol.default_params(row, **over)
    M0, C0 = initial(kind, row, col)
    o = ol.Oracle(row, col, p)
    o.set_state(M0, C0)
    out = dict(row=row, col=col, M0=M0, C0=C0, steps=steps,
               over_keys=np.array(list(over.keys()), dtype="U16"),
               over_vals=np.array(list(over.values()), dtype=np.float64))
    # the same run with the OpenMP build: env[k] = how far the reference's two
    # builds are from each other after step k+1 (its self-consistency envelope)
    o2 = ol.Oracle(row, col, p, omp=True)
    o2.set_state(M0, C0)
    counts, nits, masses, env = [], [], [], []
    near = np.zeros((steps, 16, NEAR_CAP), dtype=np.int64)
    lib = ol.load()
    for k in range(steps):
        Mk, Ck = o.get_state()
        rn, rnear = replay_step(lib, p, row, col, Mk, Ck, o.get_mnew())
        r = o.step()
        assert rn == r["nits"], (name, k, rn, r["nits"])     # the replay is the stepper
        for si, lst in enumerate(rnear[:16]):
            near[k, si, :len(lst)] = lst
        out["near"] = near
        o2.step()
        Ma, Ca = o.get_state()
        Mb, Cb = o2.get_state()
        env.append(max(np.linalg.norm(Ma - Mb) / np.linalg.norm(Ma),
                       np.linalg.norm(Ca - Cb) / np.linalg.norm(Ca)))
        out["env"] = np.array(env)
        counts.append(r["countIters"])
        nits.append((r["nits"] + [0] * 16)[:16])
        M, C = o.get_state()
        out[f"M{k + 1}"] = M
        out[f"C{k + 1}"] = C
        masses.append((M.mean(), C.mean()))
    out["countIters"] = np.array(counts)
    out["nits"] = np.array(nits)
    return out


if __name__ == "__main__":
    ol.build_oracle()
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run(name))
        print("wrote", name)
