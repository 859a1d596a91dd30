#!/usr/bin/env python
"""bench.py -- headline benchmark of the implicit time step (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

metric  : grid-cell CG-iterations/s  (= row*col * sum of CG iterations / time)
workload: 4096 x 4096 grid, shipped parameter.txt values, deterministic
          inoculation MinitialCond = 1 with height 0.9, depth 0.3 (M < 1, a few
          hundred CG iterations per solve).  A "step" is one
          implicit time step: D(M), Picard loop, CG solves, C update.
N > 1   : row slabs over N B200s (strong scaling: the grid is fixed), launched by
          torch.distributed.run, one rank per GPU; torch.distributed is used only
          to hand out the NCCL id and to take the max of the per-rank times.

One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))

METRIC = "grid-cell CG-iterations/s"
UNIT = "cell-iterations/s"
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md


def ic1(row, col, height, depth, row0=0, nrows=None):
    """MinitialCond = 1 of the reference (PDE-ODEsolver.f90:279-288), C = 1 (:276)."""
    xDel = 1.0 / row
    a = -height / ((depth * depth) * (depth * depth))
    nrows = row if nrows is None else nrows          # only this slab's rows
    t = np.arange(row0, row0 + nrows, dtype=np.float64) * xDel
    f = a * ((t * t) * (t * t)) + height
    f = np.where(f <= 0, 0.0, f)
    return np.repeat(f, col), np.ones(nrows * col)


def ic_default_like(row, col, seed=7, npts=60, height=0.02, depth=0.0390625):
    """Shipped parameter.txt inoculation (MinitialCond = 5, :328-352) with a fixed seed."""
    rng = np.random.RandomState(seed)
    xDel = 1.0 / row
    a = -height / (depth * depth)
    x = (np.arange(1, col + 1) % col) * xDel
    M = np.zeros((row, col))
    # only rows within `depth` + 0.1 of the y = 0 side can be touched
    nr = min(row, int((0.1 + depth) * row) + 2)
    y = ((np.arange(nr)[:, None] * col + np.arange(1, col + 1)[None, :]) // row) * xDel
    for _ in range(npts):
        xr, yr = rng.random_sample(), rng.random_sample() * 0.1
        innoc = a * ((x[None, :] - xr) ** 2 + (y - yr) ** 2) + height
        M[:nr] += np.where(innoc >= 0, innoc, 0.0)
    nz = M > 0
    M *= height / (M[nz].sum() / nz.sum())
    return M.ravel(), np.ones(row * col)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(row, col, height, depth, cg_iters):
    """The oracle (a port of the reference, not gfortran) on the host cores:
    the first solve of step 0 of the same workload, truncated to `cg_iters`
    CG iterations.  Returns (cell-iterations/s, cores, seconds, iterations)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    ol.build_oracle()
    p = ol.default_params(row)
    o = ol.Oracle(row, col, p, omp=True)
    M0, C0 = ic1(row, col, height, depth)
    o.set_state(M0, C0)
    o.set_maxit(cg_iters - 1)          # CG:88 stops once nit > maxit
    t0 = time.perf_counter()
    nit, _ = o.first_solve_trace(1)
    dt = time.perf_counter() - t0
    cores = o.lib.oracle_num_threads()
    o.close()
    return row * col * nit / dt, cores, dt, nit


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the image has
    no Fortran compiler) on all host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    row = col = args.grid
    iters = args.ref_cg_iters
    vals = []
    for k in range(args.warmup + args.steps):
        v, cores, dt, nit = cpu_baseline(row, col, args.height, args.depth, iters)
        if k >= args.warmup:
            vals.append((v, dt, nit))
    tot_it = sum(n for _, _, n in vals)
    tot_t = sum(t for _, t, _ in vals)
    value = row * col * tot_it / tot_t
    sample = (f"each step = operator assembly + the first {iters} CG iterations of step 0's first "
              f"solve on the same {row}x{col} IC-1 grid (oracle port of CGdiag.f90, OpenMP)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def workload_config(args, nranks):
    return {
        "workload": f"{args.grid}x{args.grid} grid, shipped parameter.txt values, IC MinitialCond=1 "
                    f"height={args.height} depth={args.depth}; one step = one implicit time step "
                    "(D(M), Picard loop, CG solves, C update)",
        "grid": [args.grid, args.grid],
        "parallelism": f"row-slabs x{nranks}" if nranks > 1 else "single GPU",
        "l2": ("inputs larger than L2: " if 12 * 8 * args.grid * args.grid > 4 * 126e6 else
               "NOT larger than L2 (small --grid; no roofline claim): ") +
              "12 fp64 fields of the grid stay resident "
              f"({12 * 8 * args.grid * args.grid / 1e6:.0f} MB total vs 126 MB L2); no flush",
    }


_REAL_STDOUT = None


def quiet_stdout():
    """Everything but the final JSON line goes to stderr: NCCL prints a version
    banner on stdout, and the contract is ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--height", type=float, default=0.9)
    ap.add_argument("--depth", type=float, default=0.3)
    ap.add_argument("--ref-cg-iters", type=int, default=24)
    ap.add_argument("--cpu-cg-iters", type=int, default=120)
    ap.add_argument("--kernel-reps", type=int, default=50)
    ap.add_argument("--cg-cap", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                      # timing rule: W >= 3

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import pdeode_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(pb.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    row = col = args.grid
    n = row * col
    params = pb.default_params(row)
    s = pb.Solver(row, col, params, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    exchange = s.exchange_mode()
    stream = torch.cuda.Stream()           # the library launches here; events are recorded here
    s.set_stream(stream.cuda_stream)
    s.set_maxit(args.cg_cap)               # safety net only; a hit invalidates the run (checked below)
    M0, C0 = ic1(row, col, args.height, args.depth, s.row0, s.nrows)
    Mh = torch.from_numpy(M0).pin_memory()
    Ch = torch.from_numpy(C0).pin_memory()
    del M0, C0
    Mout = torch.empty_like(Mh).pin_memory()
    Cout = torch.empty_like(Ch).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """fn() -> CG iterations; returns (max-over-ranks seconds, iterations)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        its = 0
        for _ in range(steps):
            its += fn()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), its

    # ---- resident-state throughput (inputs already in HBM) ----
    s.set_state_ptr(Mh.data_ptr(), Ch.data_ptr())

    def resident_step():
        r = s.step()
        if r["rc"] != 0:
            raise SystemExit(f"bench.py: pdeode_step returned {r['rc']} (CG cap or Picard cap hit)")
        return r["nitSum"]

    for _ in range(args.warmup):
        resident_step()
    l0 = s.launch_count()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    t_res, it_res = timed(resident_step, args.steps)
    clocks = clk.stop() if rank == 0 else None
    launches = s.launch_count() - l0
    value = n * it_res / t_res

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            s.set_state_ptr(Mh.data_ptr(), Ch.data_ptr())      # H2D of this step's inputs
            k = s.step()["nitSum"]
            s.get_state_ptr(Mout.data_ptr(), Cout.data_ptr())  # D2H of the result
            s.mass()
            return k
        e2e_step()
        t_e, it_e = timed(e2e_step, args.steps)
        e2e = {"value": n * it_e / t_e, "unit": UNIT,
               "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 2 * 8 * n + 16,
               "ms_per_step": 1e3 * t_e / args.steps, "cg_iterations": it_e,
               "note": "every step: pdeode_set_state (pinned host -> device), pdeode_step, "
                       "pdeode_get_state, pdeode_mass; set_state restarts the CG warm start"}

    # ---- shipped-parameter run: wall time per step ----
    default_ms = None
    if world == 1 and args.grid <= 4096:
        Md, Cd = ic_default_like(row, col)
        s.set_state(Md, Cd)
        for _ in range(3):
            s.step()
        t_d, it_d = timed(lambda: s.step()["nitSum"], 10)
        default_ms = {"ms_per_step": 1e3 * t_d / 10, "cg_iterations_per_step": it_d / 10}

    # ---- roofline of the dominant kernel, timed live with CUDA events ----
    roof = None
    if world == 1:
        s.set_state_ptr(Mh.data_ptr(), Ch.data_ptr())
        s.step()
        peak, peak_src = measured_peak()
        ks = {}
        info = s.engine_info()
        deferred = info.get("defer_x") == "1"
        # 88-byte iteration: x += alfa p rides in the SpMV kernel, the update kernel does r only
        kinds = ((("cg_spmv", pb.KERNEL_CG_SPMV_X), ("cg_update", pb.KERNEL_CG_UPDATE_R)) if deferred
                 else (("cg_spmv", pb.KERNEL_CG_SPMV), ("cg_update", pb.KERNEL_CG_UPDATE)))
        for name, kid in kinds:
            ms = s.time_kernel(kid, args.kernel_reps)
            by = pb.Solver.bytes_per_cell(kid) * n
            ks[name] = {"ms": ms, "GBps": by / ms / 1e6, "bytes_per_cell": pb.Solver.bytes_per_cell(kid)}
        dom = max(ks, key=lambda k: ks[k]["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get(dom + ({"cg_spmv": "_x", "cg_update": "_r"}[dom] if deferred else ""))
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": ks[dom]["GBps"], "peak": peak,
                "unit": "GB/s", "frac": ks[dom]["GBps"] / peak, "traffic": traffic,
                "peak_source": peak_src, "kernels": ks, "engine": info,
                "cg_iteration": {
                    "bytes_per_cell": 96,
                    "bytes_per_cell_moved": sum(k["bytes_per_cell"] for k in ks.values()),
                    "ms": ks["cg_spmv"]["ms"] + ks["cg_update"]["ms"],
                    "GBps": 96 * n / (ks["cg_spmv"]["ms"] + ks["cg_update"]["ms"]) / 1e6,
                    "frac": 96 * n / (ks["cg_spmv"]["ms"] + ks["cg_update"]["ms"]) / 1e6 / peak}}
    s.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dt, nit = cpu_baseline(row, col, args.height, args.depth, args.cpu_cg_iters)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"operator assembly + first {nit} CG iterations of step 0's first solve, same "
                         f"{row}x{col} grid and IC, oracle port with OpenMP ({dt:.1f} s)"}

    if rank == 0:
        cfg = workload_config(args, world)
        cfg["exchange"] = {"single": "none (one GPU)",
                           "nccl": "NCCL send/recv of one ghost row + 2 all-reduces per CG iteration",
                           "peer": "inside K_A/K_B over NVLink peer memory (CUDA IPC): ghost rows and "
                                   "dot-product slots written straight into the other GPUs"}[exchange]
        if default_ms:
            cfg["shipped_parameters_run"] = default_ms
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cg_iterations": it_res, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        }
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
