#!/usr/bin/env python
"""Row-slab multi-GPU check (GPU box, under torch.distributed.run).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [--grid 1024] [--steps 3]

Every rank steps its slab of an N-rank solve; rank 0 also runs the same problem
on one GPU.  Checks: identical Picard counts, CG counts within +-2, gathered
fields within the parity tolerance, identical mass / peak diagnostics.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))
sys.path.insert(0, ROOT)
import pdeode_b200 as pb  # noqa: E402
from bench import ic1  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(pb.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(pb.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    nccl_id = idt.cpu().numpy().tobytes()

    row = args.grid
    col = args.cols or args.grid
    p = pb.default_params(row)
    M0, C0 = ic1(row, col, 0.9, 0.3)
    rng = np.random.RandomState(3)
    M0 = M0 * (1 + 0.02 * rng.rand(M0.size))          # break the column symmetry
    s = pb.Solver(row, col, p, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    lo, hi = s.row0 * col, (s.row0 + s.nrows) * col
    s.set_state(M0[lo:hi], C0[lo:hi])
    ref = None
    if rank == 0:
        ref = pb.Solver(row, col, p, device=local)
        ref.set_state(M0, C0)
    ok = True
    if rank == 0:
        print("exchange mode:", s.exchange_mode(), flush=True)
    for k in range(args.steps):
        g = s.step_trace()
        mass = s.mass()
        peak = s.peak(-1.0, -1.0, -1.0)
        Ml, Cl = s.get_state()
        parts = [torch.zeros(0)] * world
        sizes = [((r + 1) * row // world - r * row // world) * col for r in range(world)]
        Mg = [torch.empty(sz, dtype=torch.float64, device="cuda") for sz in sizes]
        Cg = [torch.empty(sz, dtype=torch.float64, device="cuda") for sz in sizes]
        dist.all_gather(Mg, torch.from_numpy(Ml).cuda())
        dist.all_gather(Cg, torch.from_numpy(Cl).cuda())
        if rank == 0:
            r1 = ref.step_trace()
            M1, C1 = ref.get_state()
            Mall = torch.cat(Mg).cpu().numpy(); Call = torch.cat(Cg).cpu().numpy()
            eM = np.linalg.norm(Mall - M1) / np.linalg.norm(M1)
            eC = np.linalg.norm(Call - C1) / np.linalg.norm(C1)
            same = g["nits"] == r1["nits"]
            tol = 1e-9 if same else 2e-7
            good = (g["countIters"] == r1["countIters"]
                    and all(abs(a - b) <= 2 for a, b in zip(g["nits"], r1["nits"]))
                    and eM <= tol and eC <= tol)
            m1 = ref.mass(); p1 = ref.peak(-1.0, -1.0, -1.0)
            good = good and abs(mass[0] - m1[0]) <= 1e-9 * abs(m1[0]) and abs(mass[1] - m1[1]) <= 1e-9
            good = good and abs(peak[1] - p1[1]) <= 1e-7 and peak[2] == p1[2]
            print(f"step {k}: N={world} nits {g['nits']} vs 1-GPU {r1['nits']}  relL2 M {eM:.2e} C {eC:.2e} "
                  f"mass {mass} vs {m1} peak {peak} vs {p1} -> {'ok' if good else 'FAIL'}", flush=True)
            ok = ok and good
    # initial conditions built on the device, slab by slab, and decimated frames
    rs = np.random.RandomState(1)
    px, py = rs.rand(9), rs.rand(9) * 0.1
    for ic in (5, 9, 1):
        s.set_initial_condition(ic, 0.2, 0.3, px, py)
        Ml, Cl = s.get_state()
        fM, fC = s.frame(4)
        sizes = [((r + 1) * row // world - r * row // world) * col for r in range(world)]
        Mg = [torch.empty(sz, dtype=torch.float64, device="cuda") for sz in sizes]
        dist.all_gather(Mg, torch.from_numpy(Ml).cuda())
        fr = [None] * world
        dist.all_gather_object(fr, fM)
        if rank == 0:
            ref.set_initial_condition(ic, 0.2, 0.3, px, py)
            M1, _ = ref.get_state()
            Mall = torch.cat(Mg).cpu().numpy()
            same = np.array_equal(Mall, M1) if ic != 5 else np.allclose(Mall, M1, rtol=1e-14, atol=0)
            f1, _ = ref.frame(4)
            fall = np.concatenate(fr, axis=0)
            fsame = fall.shape == f1.shape and (np.array_equal(fall, f1) if ic != 5 else np.allclose(fall, f1, rtol=1e-14))
            print(f"device IC {ic}: fields {'ok' if same else 'FAIL'}, frames {'ok' if fsame else 'FAIL'}", flush=True)
            ok = ok and same and fsame
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST CHECK", "PASSED" if ok else "FAILED", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
