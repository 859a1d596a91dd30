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

One JSON line is printed by rank 0.  Besides the contract's keys it carries
  parity        (N > 1) the slabs gathered on rank 0 against the same steps on a 1-rank context:
                relative L2 of M and C, CG and Picard totals of both; a failure exits non-zero
  strong_16384  the north-star strong-scaling point measured in the same run: 16384 x 16384 at
                the shipped exponents and at the stiff ones (beta = 6, height 0.95), so that
                T1 / (N * TN) can be computed from two lines of this program
  --impl reference times the reference's own CPU code (oracle/_ref/libref_omp.so: its Fortran
  sources translated to C where they lie; else the oracle port) on ALL host cores: the OpenMP
  thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1).
"""
import argparse
import json
import os
import shutil
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
    """nvidia-smi clocks / throttle reasons during the timed region.

    start() is called BEFORE the warm-up steps and begin() right before the timed region: the
    sampler's own start-up (process spawn, NVML attaching to every GPU of the box: some hundred
    milliseconds during which launches and context operations stall -- it doubled the wall time
    of the ensemble sweep, see clock_snapshot) then falls into the warm-up, and only the samples
    stamped inside [begin, stop] are reported."""
    FIELDS = ("timestamp", "index", "clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.active",
              "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap")
    Q = ",".join(FIELDS)
    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index = index
        self.period_ms = period_ms
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t_begin = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def begin(self):
        self.t_begin = time.time()

    @staticmethod
    def parse_line(line):
        """-> (unix time or None, sm MHz, max sm MHz, [reasons]) or None for a line that is not a sample."""
        c = [t.strip() for t in line.split(",")]
        if len(c) < len(ClockSampler.FIELDS):
            return None
        try:
            sm, mx = float(c[2]), float(c[3])
        except ValueError:
            return None
        ts = None
        try:                                     # "2026/10/20 13:22:01.123", local time
            head, _, ms = c[0].partition(".")
            ts = time.mktime(time.strptime(head, "%Y/%m/%d %H:%M:%S")) + (float("0." + ms) if ms.isdigit() else 0.0)
        except (ValueError, OverflowError):
            ts = None
        reasons = [n for n, v in zip(ClockSampler.REASONS, c[6:10]) if v.lower().startswith("active")]
        return ts, sm, mx, reasons

    def stop(self):
        t_end = time.time()
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = [r for r in (self.parse_line(l) for l in self.f) if r is not None]
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        window = "all samples since the sampler started"
        if self.t_begin is not None:
            slack = 1e-3 * self.period_ms
            inside = [r for r in rows if r[0] is not None and self.t_begin - slack <= r[0] <= t_end + slack]
            if inside:
                rows, window = inside, "samples stamped inside the timed region (sampler started before the warm-up)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "reasons": sorted({n for r in rows for n in r[3]}), "samples": len(rows), "window": window}


def clock_snapshot(index):
    """One nvidia-smi query (clocks, throttle reasons) -- for runs a sampler would disturb."""
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
           "when": "one query right after the sweep (no sampler during it)"}
    try:
        txt = subprocess.run(["nvidia-smi", "-i", str(index), f"--query-gpu={ClockSampler.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=30).stdout
        r = ClockSampler.parse_line(txt.strip().splitlines()[0])
        out["sm_mhz"], out["sm_max_mhz"], out["samples"], out["reasons"] = r[1], r[2], 1, sorted(r[3])
    except Exception:
        pass
    return out


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def force_omp_threads():
    """All host cores for the CPU arm, whatever the launcher exported (torch.distributed.run
    sets OMP_NUM_THREADS=1 in every rank's environment).  Returns the thread count set."""
    import ctypes
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        g = ctypes.CDLL("libgomp.so.1")      # the runtime liboracle_omp.so / libref_omp.so link to
        g.omp_set_num_threads(n)
        g.omp_get_max_threads.restype = ctypes.c_int
        return int(g.omp_get_max_threads())
    except OSError:
        return n


def reference_cpu(row, col, height, depth, cg_iters):
    """The reference's own CPU code on the host cores: GenMatrixM + solveLSDIAG
    (PDE-ODEsolver.f90:517-590, CGdiag.f90:1-95) of step 0's first Picard iterate, the solve
    truncated to `cg_iters` iterations through its own iteration cap (CG:88).
    oracle/_ref/libref_omp.so is the reference's Fortran translated to C where it lies (no
    Fortran compiler in this image) and built with the flags of runAll_paral.sh:6; without it
    the oracle port is timed.  Returns (cell-iterations/s, cores, seconds, iterations, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cores = force_omp_threads()
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_omp.so")
    if not os.path.exists(ref_so):
        v, c2, dt, nit = cpu_baseline(row, col, height, depth, cg_iters)
        return v, c2, dt, nit, "port"
    import oracle_lib as ol
    import ref_lib as rl
    force_omp_threads()
    ref = rl.Ref(omp=True)
    cores = force_omp_threads()
    p = ol.default_params(row)
    M0, C0 = ic1(row, col, height, depth)
    n = row * col
    t0 = time.perf_counter()
    A, ioff, rhs = ref.gen_matrix(M0, C0, row, col, p)
    sol, nit, _, _, stop = ref.solve_ls_diag(A, ioff, np.zeros(n), rhs, cg_iters - 1, p.e1, p.e2)
    dt = time.perf_counter() - t0
    return n * nit / dt, cores, dt, nit, "reference"


def cpu_baseline(row, col, height, depth, cg_iters):
    """The oracle (a port of the reference, not gfortran) on the host cores:
    the first solve of step 0 of the same workload, truncated to `cg_iters`
    CG iterations.  Returns (cell-iterations/s, cores, seconds, iterations)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    force_omp_threads()
    import oracle_lib as ol
    ol.build_oracle()
    force_omp_threads()
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
    """--impl reference: the reference's own CPU implementation on ALL host cores, same config /
    metric / unit as the b200 arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    row = col = args.grid
    iters = args.ref_cg_iters
    vals = []
    for k in range(args.warmup + args.steps):
        v, cores, dt, nit, kind = reference_cpu(row, col, args.height, args.depth, iters)
        if k >= args.warmup:
            vals.append((v, dt, nit))
    tot_it = sum(n for _, _, n in vals)
    tot_t = sum(t for _, t, _ in vals)
    value = row * col * tot_it / tot_t
    what = ("GenMatrixM + solveLSDIAG of the reference (its .f90 sources translated to C by "
            "oracle/f90c and built with -O3 -fopenmp, as runAll_paral.sh:6)" if kind == "reference"
            else "the oracle port of GenMatrixM + CGdiag.f90 (OpenMP)")
    sample = (f"each step = operator assembly + the first {iters} CG iterations of step 0's first "
              f"solve on the same {row}x{col} IC-1 grid: {what}")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample, "host_cores": host_threads()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if cores == 1 and host_threads() > 1:
        out["cpu_baseline"]["warning"] = "ran on one thread of a multi-core host"
    emit(out)


def run_ensemble(args):
    """BASELINE config 5: `--ensemble` independent solves (the reference's sweep was numbered
    parameter files, .gitignore:34-35; one <name>Output/ each as runAll.sh:13-20) dealt over
    the first --gpus GPUs by pde-odesolver_b200/bin/pdeode_ensemble -- replicas, no
    communication.  Wall clock around the whole program: contexts, files and output included."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ensemble_bench as eb
    exe = os.path.join(ROOT, "pde-odesolver_b200", "bin", "pdeode_ensemble")
    grid, steps, members = args.ensemble_grid, args.ensemble_steps, args.ensemble
    n = grid * grid
    # every member writes the reference's text frames (512^2: two frames of 262 144 "x y M C"
    # lines, 54 MB per member, 13.9 GB for 256 members): on the box's disk that is the whole wall
    # time (1.1 GB/s), so the run directory goes to RAM-backed /dev/shm when it has the room
    where = args.ensemble_dir
    if where is None:
        need = members * 2 * (grid // max(1, (grid - 1) // 256 if grid > 257 else 1)) ** 2 * 110 * 1.2
        try:
            st = os.statvfs("/dev/shm")
            where = "/dev/shm" if st.f_bavail * st.f_frsize > need else None
        except OSError:
            where = None
    with tempfile.TemporaryDirectory(dir=where) as d:
        files = []
        for k in range(1, members + 1):
            with open(os.path.join(d, f"{k}.txt"), "w") as fh:
                fh.write(eb.TEMPLATE.format(**eb.member_params(k, grid, steps)))
            files.append(f"{k}.txt")
        env = dict(os.environ, PDEODE_MEMBERS_PER_GPU=str(args.ensemble_per_gpu), PDEODE_SEED="1",
                   PDEODE_ENSEMBLE_TIMING="1",
                   PDEODE_DEVICES=",".join(str(i) for i in range(max(1, args.gpus))))
        # warm-up: one short member per GPU (a fresh box pages the driver and the library in on
        # first use: 6 s of a cold 256-member sweep, measured), untimed, then its output removed
        wd = os.path.join(d, "warmup")
        os.makedirs(wd)
        wfiles = []
        for k in range(1, max(1, args.gpus) + 1):
            with open(os.path.join(wd, f"{k}.txt"), "w") as fh:
                fh.write(eb.TEMPLATE.format(**eb.member_params(k, grid, 20)))
            wfiles.append(f"{k}.txt")
        w = subprocess.run([exe] + wfiles, cwd=wd, env=dict(env, PDEODE_MEMBERS_PER_GPU="1"),
                           capture_output=True, text=True)
        if w.returncode != 0:
            raise SystemExit("bench.py --ensemble: warm-up sweep failed: " + (w.stdout + w.stderr)[-800:])
        shutil.rmtree(wd)
        log("warm-up sweep done")
        clk = ClockSampler(0, args.ensemble_clock_ms)
        if args.ensemble_clock_ms > 0:
            clk.start()
        t0 = time.perf_counter()
        r = subprocess.run([exe] + files, cwd=d, env=env, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if args.ensemble_clock_ms > 0:
            clocks = clk.stop()
            clocks["when"] = f"sampled every {args.ensemble_clock_ms} ms during the sweep"
        else:
            clocks = clock_snapshot(0)
        if r.returncode != 0:
            raise SystemExit("bench.py --ensemble: pdeode_ensemble failed: " + (r.stdout + r.stderr)[-800:])
        breakdown = [l for l in r.stderr.splitlines() if l.startswith("pdeode_ensemble timing")]
        log(f"pdeode_ensemble: {dt:.2f} s wall; " + (breakdown[0] if breakdown else ""))
        cg = 0.0
        solver_s = []
        out_bytes = 0
        for k in range(1, members + 1):
            for fn in os.listdir(os.path.join(d, f"{k}Output")):
                out_bytes += os.path.getsize(os.path.join(d, f"{k}Output", fn))
            txt = open(os.path.join(d, f"{k}Output", "statReport.dat")).read()
            cg += eb.stat_value(txt, "Avg Iters for linear") * eb.stat_value(txt, "Avg Iters for iterating") * steps
            solver_s.append(eb.stat_value(txt, "Time to compute"))
    emit({
        "metric": METRIC, "value": n * cg / dt, "unit": UNIT, "n_gpus": max(1, args.gpus),
        "steps": steps, "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"ensemble: {members} independent {grid}x{grid} solves of {steps} time steps "
                               "(numbered parameter files varying delta, nu, kappa, gama, beta, height), "
                               f"{args.ensemble_per_gpu} members in flight per GPU, replicas only",
                   "grid": [grid, grid], "members": members, "parallelism": f"replicas over {max(1, args.gpus)} GPUs"},
        "wall_s": dt, "cg_iterations": round(cg), "clocks": clocks,
        "program_timing": breakdown[0] if breakdown else None,
        "output_bytes": out_bytes, "output_dir": where or tempfile.gettempdir(),
        "member_time_to_compute_s": {"median": float(np.median(solver_s)), "max": float(max(solver_s))},
        "gpu_launches": None,
        "note": "wall clock of the whole pdeode_ensemble program, start to exit: context creation, "
                "console and output files of every member included; warmup = one untimed sweep of one "
                "20-step member per GPU",
    })


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


_T0 = time.perf_counter()


def log(msg):
    """Progress on stderr (the one JSON line owns stdout)."""
    sys.stderr.write(f"[bench {time.perf_counter() - _T0:7.1f} s] {msg}\n")
    sys.stderr.flush()


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
    ap.add_argument("--no-strong16384", action="store_true",
                    help="skip the 16384^2 strong-scaling sub-record")
    ap.add_argument("--strong-cg-cap", type=int, default=600,
                    help="CG iterations per step of the stiff 16384^2 sub-record")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the check against one rank")
    ap.add_argument("--ensemble", type=int, default=0,
                    help="BASELINE config 5 instead of the headline workload: this many independent "
                         "512^2 solves (numbered parameter files) spread over --gpus GPUs by "
                         "pdeode_ensemble, one line with the aggregate throughput")
    ap.add_argument("--ensemble-grid", type=int, default=512)
    ap.add_argument("--ensemble-steps", type=int, default=200)
    ap.add_argument("--ensemble-per-gpu", type=int, default=16)
    ap.add_argument("--ensemble-clock-ms", type=int, default=0,
                    help="nvidia-smi sampling period during the sweep; 0 (default): one query right after "
                         "it instead -- the sweep is bound by launch and context latencies, and a sampler "
                         "running next to it more than doubles its wall time (measured: 32 members on one "
                         "B200 2.23 s without, 4.5-5.4 s with nvidia-smi -lms 100 / 1000)")
    ap.add_argument("--ensemble-dir", default=None,
                    help="where the members' run directories go (default: /dev/shm if it has room, else the temp dir)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                      # timing rule: W >= 3

    if args.impl == "reference":
        run_reference(args)
        return
    if args.ensemble > 0:
        run_ensemble(args)
        return

    if int(os.environ.get("WORLD_SIZE", "1")) == 1:
        force_omp_threads()                  # the cpu_baseline leg below uses all host cores
    import torch
    import pdeode_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_nccl_id():
        """One NCCL id per multi-rank context, made on rank 0 and handed out with torch.distributed."""
        if world == 1:
            return None
        idt = torch.zeros(pb.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    nccl_id = fresh_nccl_id()

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
    if rank == 0:
        log(f"context and host fields ready ({row}x{col}, {world} rank(s), exchange {exchange})")
    s.set_state_ptr(Mh.data_ptr(), Ch.data_ptr())

    pic_total = [0]

    def resident_step():
        r = s.step()
        if r["rc"] != 0:
            raise SystemExit(f"bench.py: pdeode_step returned {r['rc']} (CG cap or Picard cap hit)")
        pic_total[0] += r["countIters"]
        return r["nitSum"]

    clk = ClockSampler(local)
    if rank == 0:
        clk.start()                          # its start-up falls into the warm-up
    it_warm = 0
    for _ in range(args.warmup):
        it_warm += resident_step()
    l0 = s.launch_count()
    clk.begin()
    t_res, it_res = timed(resident_step, args.steps)
    clocks = clk.stop()
    clocks = clocks if rank == 0 else None
    launches = s.launch_count() - l0
    value = n * it_res / t_res
    if rank == 0:
        log(f"resident: {args.steps} steps, {it_res} CG iterations, {1e3 * t_res / args.steps:.2f} ms per step, {value:.4g} {UNIT}")

    # ---- N > 1: the slabs, gathered on rank 0, against the same steps on ONE rank ----
    parity = None
    if world > 1 and not args.no_parity:
        nsteps_done = args.warmup + args.steps
        Ml = np.empty(s.n_local); Cl = np.empty(s.n_local)
        s.get_state(Ml, Cl)
        tot = torch.tensor([it_warm + it_res, pic_total[0]], device="cuda", dtype=torch.float64)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((s.row0, Ml, Cl), gathered, dst=0)
        if rank == 0:
            gathered.sort(key=lambda t: t[0])
            Mn = np.concatenate([g[1] for g in gathered]); Cn = np.concatenate([g[2] for g in gathered])
            one = pb.Solver(row, col, params, device=local)
            one.set_maxit(args.cg_cap)
            M1, C1 = ic1(row, col, args.height, args.depth)
            one.set_state(M1, C1)
            it1 = pic1 = 0
            for _ in range(nsteps_done):
                r = one.step()
                it1 += r["nitSum"]; pic1 += r["countIters"]
            M1, C1 = one.get_state()
            one.close()
            eM = float(np.linalg.norm(Mn - M1) / np.linalg.norm(M1))
            eC = float(np.linalg.norm(Cn - C1) / np.linalg.norm(C1))
            itN, picN = int(tot[0].item()), int(tot[1].item())
            # the summation order differs between N ranks and one (rank order on top of the CTA
            # order), so CG counts may move by a few per solve and the fields by one CG step
            # (20 e2), as between the reference's own serial and OpenMP builds; Picard is exact
            ok = (picN == pic1 and abs(itN - it1) <= 2 * pic1 and eM <= 20 * params.e2 and eC <= 20 * params.e2)
            parity = {"ok": bool(ok), "steps": nsteps_done, "relL2_M": eM, "relL2_C": eC,
                      "cg_iterations": [itN, it1], "picard_iterates": [picN, pic1],
                      "bar": "Picard exact, |dCG| <= 2 per solve, relL2 <= 20*e2 = %.1e" % (20 * params.e2),
                      "against": "the same %d steps on a 1-rank context on rank 0's GPU" % nsteps_done}
        barrier()

    if rank == 0 and parity is not None:
        log(f"parity against one rank: {parity}")
    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            s.set_state_ptr(Mh.data_ptr(), Ch.data_ptr())      # H2D of this step's inputs
            k = s.step()["nitSum"]
            s.get_state_ptr(Mout.data_ptr(), Cout.data_ptr())  # D2H of the result
            s.mass()
            return k
        for _ in range(3):
            e2e_step()
        t_e, it_e = timed(e2e_step, args.steps)
        e2e = {"value": n * it_e / t_e, "unit": UNIT,
               "h2d_bytes_per_step": 2 * 8 * n, "d2h_bytes_per_step": 2 * 8 * n + 16,
               "ms_per_step": 1e3 * t_e / args.steps, "cg_iterations": it_e,
               "note": "every step: pdeode_set_state (pinned host -> device), pdeode_step, "
                       "pdeode_get_state, pdeode_mass; set_state restarts the CG warm start"}

    if rank == 0 and e2e is not None:
        log(f"e2e: {e2e['ms_per_step']:.2f} ms per step, {e2e['value']:.4g} {UNIT}")
    # ---- shipped-parameter run: wall time per step ----
    default_ms = None
    if world == 1 and args.grid <= 4096:
        Md, Cd = ic_default_like(row, col)
        s.set_state(Md, Cd)
        for _ in range(3):
            s.step()
        pic_d = [0]

        def shipped_step():
            r = s.step()
            pic_d[0] += r["countIters"]
            return r["nitSum"]

        t_d, it_d = timed(shipped_step, 10)
        # algorithmic bytes of a whole time step, SURVEY 8(d)'s per-kernel figures: D(M) once
        # (16 B per cell), per Picard iterate centre diagonal + first residual (56) and C update +
        # residuals (48), per CG iteration 96
        by_d = 16.0 + 104.0 * pic_d[0] / 10 + 96.0 * it_d / 10
        pk_d, _ = measured_peak()
        default_ms = {"ms_per_step": 1e3 * t_d / 10, "cg_iterations_per_step": it_d / 10,
                      "picard_iterates_per_step": pic_d[0] / 10,
                      "bytes_per_cell_algorithmic": by_d,
                      "GBps": by_d * n / (t_d / 10) / 1e9,
                      "frac_of_measured_hbm_peak": by_d * n / (t_d / 10) / 1e9 / pk_d}

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
    del Mh, Ch, Mout, Cout
    if rank == 0:
        log(f"shipped parameters: {default_ms}; roofline: {None if roof is None else {k: v for k, v in roof.items() if k in ('kernel', 'achieved', 'frac')}}")

    # ---- the north-star strong-scaling point, in the same run: 16384 x 16384 over `world` GPUs ----
    strong = None
    if not args.no_strong16384:
        strong = {"grid": [16384, 16384], "n_gpus": world, "unit": UNIT,
                  "note": "strong scaling: the same grid and the same CG iterations at every N; "
                          "efficiency(N) = value(N) / (N * value(1)) from two lines of this program"}
        g16 = 16384
        n16 = g16 * g16
        runs = (
            # shipped exponents, IC 1 with height 0.9: whole steps (about 2000 CG iterations each)
            ("shipped_beta4", dict(), 0.9, None, 3, 3),
            # SURVEY config 4, the stiff regime of README.md:42-44 / P:471: beta = 6, height 0.95.
            # Its solves run for 1e4 .. 1e6 iterations, so a step is cut to its first Picard
            # iterate (pdeode_set_picard_cap(0): the truncated iterate is far from converged --
            # mean |M - Mnew| of order 100 at this size -- and the Picard loop would otherwise
            # go on for hours) and --strong-cg-cap CG iterations (CG:88) of the run's first
            # step, rebuilt on the device every time: the same iterations at every N
            ("stiff_beta6", dict(beta=6), 0.95, args.strong_cg_cap, 3, 3),
        )
        for name, over, height, cap, w16, k16 in runs:
            p16 = pb.default_params(g16, **over)
            s16 = pb.Solver(g16, g16, p16, device=local, rank=rank, nranks=world, nccl_id=fresh_nccl_id())
            s16.set_stream(stream.cuda_stream)
            s16.set_maxit(cap if cap is not None else 10 * args.cg_cap)
            if cap is not None:
                s16.set_picard_cap(0)
            s16.set_initial_condition(1, args.depth, height)      # built on the device (P:279-288)

            def step16():
                if cap is not None:
                    # a truncated stiff solve leaves an iterate that is no state to continue
                    # from: every step is the first step, from fields rebuilt on the device
                    s16.set_initial_condition(1, args.depth, height)
                return s16.step()["nitSum"]

            clk16 = ClockSampler(local)
            if rank == 0:
                clk16.start()
            for _ in range(w16):
                step16()
            clk16.begin()
            t16, it16 = timed(step16, k16)
            c16 = clk16.stop()
            c16 = c16 if rank == 0 else None
            mM, mC = s16.mass()
            strong[name] = {"value": n16 * it16 / t16, "ms_per_step": 1e3 * t16 / k16, "steps": k16,
                            "warmup": w16, "cg_iterations": it16, "clocks": c16,
                            "params": dict(over, height=height, cg_cap_per_solve=cap),
                            "mean_M": mM, "mean_C": mC, "engine": s16.engine_info(),
                            "exchange": s16.exchange_mode()}
            if not (np.isfinite(mM) and np.isfinite(mC)):
                strong[name]["invalid"] = "non-finite state"
            s16.close()
            if rank == 0:
                log(f"16384^2 {name}: {strong[name]['ms_per_step']:.1f} ms per step, {it16} CG iterations, {strong[name]['value']:.4g} {UNIT}")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, dt, nit, kind = reference_cpu(row, col, args.height, args.depth, args.cpu_cg_iters)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "host_cores": host_threads(),
               "sample": f"operator assembly + first {nit} CG iterations of step 0's first solve, same "
                         f"{row}x{col} grid and IC ({dt:.1f} s): " +
                         ("the reference's GenMatrixM + solveLSDIAG (its .f90 translated to C by oracle/f90c, "
                          "-O3 -fopenmp as runAll_paral.sh:6)" if kind == "reference" else
                          "oracle port with OpenMP")}

    if rank == 0 and cpu is not None:
        log(f"cpu baseline: {cpu['value']:.4g} {UNIT} on {cpu['cores']} threads ({cpu['kind']})")
    if rank == 0:
        cfg = workload_config(args, world)         # the same dict the reference arm prints
        exchange_txt = {"single": "none (one GPU)",
                        "nccl": "NCCL send/recv of one ghost row + 2 all-reduces per CG iteration",
                        "peer": "inside K_A/K_B over NVLink peer memory (CUDA IPC): ghost rows and "
                                "dot-product slots written straight into the other GPUs"}[exchange]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cg_iterations": it_res, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "exchange": exchange_txt, "shipped_parameters_run": default_ms,
            "parity": parity, "strong_16384": strong,
        }
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the %d-rank run does not match one rank: %r" % (world, parity))


if __name__ == "__main__":
    main()
