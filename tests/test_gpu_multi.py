"""Row-slab multi-GPU parity, driver-visible: tools/dist_check.py under torch.distributed.run
on N = 2 / 4 / 8 of the visible GPUs, for the exchange inside the CG kernels over NVLink peer
memory (PDEODE_PEER=1, default) and for the NCCL send/recv + all-reduce path (PDEODE_PEER=0),
with the kernel-per-phase engine and the persistent solve.  dist_check steps every slab,
gathers them and compares with the same steps on one GPU: Picard counts exact, CG counts within
+-2 per solve, fields to 1e-9 (2e-7 once a count differs), mass / peak diagnostics, device-built
initial conditions and decimated frames.  SKIPPED -- not passed -- where fewer GPUs are visible.
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_dist_check(n, env_over, grid, cols=0, steps=2):
    env = dict(os.environ)
    for k in ("PDEODE_PEER", "PDEODE_PERSIST", "PDEODE_WHOLE_STEP", "PDEODE_SUM_ORDER"):
        env.pop(k, None)
    env.update(env_over)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tools", "dist_check.py"), "--grid", str(grid), "--steps", str(steps)]
    if cols:
        cmd += ["--cols", str(cols)]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("peer,persist", [("1", "1"), ("1", "0"), ("0", "0")])
def test_slabs_match_one_gpu(n, peer, persist):
    if gpu_count() < n:
        pytest.skip(f"{n} GPUs needed, {gpu_count()} visible")
    rc, out = run_dist_check(n, {"PDEODE_PEER": peer, "PDEODE_PERSIST": persist}, 1024)
    want = "peer" if peer == "1" else "nccl"
    assert f"exchange mode: {want}" in out, out[-3000:]
    assert rc == 0 and "DIST CHECK PASSED" in out, out[-3000:]


@pytest.mark.parametrize("n", [2, 4])
def test_uneven_slabs_and_narrow_grid(n):
    """Rows not divisible by N, columns not a multiple of the 16-double pitch."""
    if gpu_count() < n:
        pytest.skip(f"{n} GPUs needed, {gpu_count()} visible")
    rc, out = run_dist_check(n, {}, 1030, cols=700)
    assert rc == 0 and "DIST CHECK PASSED" in out, out[-3000:]


def test_second_device_after_first():
    """One process, two contexts on different devices, both on the kernel-per-phase engine with
    the bulk-async SpMV (> 80 KB of dynamic shared memory: the attribute is per device)."""
    if gpu_count() < 2:
        pytest.skip("2 GPUs needed")
    code = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(%r, "pde-odesolver_b200", "python"))
sys.path.insert(0, os.path.join(%r, "tests"))
os.environ["PDEODE_PERSIST"] = "0"; os.environ["PDEODE_WHOLE_STEP"] = "0"
import pdeode_b200 as pb
import oracle_lib as ol
row = col = 2304                      # > 4M cells: k_spmv_bulk<256,4>
M0, C0 = ol.ic1(row, col, 0.9, 0.3)
out = []
for dev in (0, 1):
    s = pb.Solver(row, col, pb.default_params(row), device=dev)
    s.set_state(M0, C0)
    r = s.step_trace()
    out.append((r["nits"], s.get_state()))
    s.close()
assert out[0][0] == out[1][0], (out[0][0], out[1][0])
assert np.array_equal(out[0][1][0], out[1][1][0]) and np.array_equal(out[0][1][1], out[1][1][1])
print("TWO DEVICES OK", out[0][0])
''' % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "TWO DEVICES OK" in r.stdout, (r.stdout + r.stderr)[-3000:]


# ------------------------------------------------------------ several GPUs behind ONE process

def _group_case(pb, n, row, col, steps, env=None):
    import numpy as np
    import oracle_lib as ol
    M0, C0 = ol.ic1(row, col, 0.9, 0.3)
    rng = np.random.RandomState(3)
    M0 = M0 * (1 + 0.02 * rng.rand(M0.size))
    p = pb.default_params(row)
    one = pb.Solver(row, col, p, device=0)
    grp = pb.Solver(row, col, p, ngpus=n)
    assert grp.exchange_mode() in ("peer", "nccl") and grp.nrows == row and grp.row0 == 0
    one.set_state(M0, C0); grp.set_state(M0, C0)
    for k in range(steps):
        a = one.step_trace(); b = grp.step_trace()
        assert a["countIters"] == b["countIters"], (k, a, b)
        assert all(abs(x - y) <= 2 for x, y in zip(a["nits"], b["nits"])), (k, a["nits"], b["nits"])
        Ma, Ca = one.get_state(); Mb, Cb = grp.get_state()
        tol = 1e-9 if a["nits"] == b["nits"] else 2e-7
        assert np.linalg.norm(Ma - Mb) <= tol * np.linalg.norm(Ma), k
        assert np.linalg.norm(Ca - Cb) <= tol * np.linalg.norm(Ca), k
        ma, mb = one.mass(), grp.mass()
        assert abs(ma[0] - mb[0]) <= 1e-9 * abs(ma[0]) and abs(ma[1] - mb[1]) <= 1e-9
        assert one.peak(-1.0, -1.0, -1.0)[2] == grp.peak(-1.0, -1.0, -1.0)[2]
    fa, fb = one.frame(4), grp.frame(4)
    assert fa[0].shape == fb[0].shape
    assert np.allclose(fa[0], fb[0], rtol=0, atol=2e-7) and np.allclose(fa[1], fb[1], rtol=0, atol=2e-7)
    # fields built on the device, slab by slab: the same bits as on one GPU
    one.set_initial_condition(9, 0.2, 0.3, np.linspace(0.1, 0.9, 5), np.zeros(5))
    grp.set_initial_condition(9, 0.2, 0.3, np.linspace(0.1, 0.9, 5), np.zeros(5))
    assert np.array_equal(one.get_state()[0], grp.get_state()[0])
    sm = grp.step_many(3); so = one.step_many(3)
    assert sm["stepsDone"] == so["stepsDone"] == 3 and sm["countSum"] == so["countSum"]
    one.close(); grp.close()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_group_context_matches_one_gpu(pb, n):
    """pdeode_create_group: N GPUs driven from one process (one worker thread per GPU, peer
    access instead of CUDA IPC) behave like one GPU holding the whole grid."""
    if gpu_count() < n:
        pytest.skip(f"{n} GPUs needed, {gpu_count()} visible")
    _group_case(pb, n, 1030, 700, 2)


def test_drop_in_program_on_several_gpus(tmp_path):
    """`PDEODE_NGPUS=N echo FILE | pdeode_solver` (what runAll.sh:7 runs as ./a.out) writes the
    same output.dat / total.dat as on one GPU."""
    n = min(gpu_count(), 8)
    if n < 2:
        pytest.skip("2 GPUs needed")
    import numpy as np
    import __graft_entry__ as ge
    ge.build_host()
    exe = os.path.join(ROOT, "pde-odesolver_b200", "bin", "pdeode_solver")
    # the shipped parameters on 1024^2, cut to 41 steps and 4 frames (the file is positional)
    src = open(os.path.join(ROOT, "tests", "data", "shipped1024.txt")).read()
    src = src.replace("nOuts     = 90", "nOuts     = 4").replace("tEnd      = 30", "tEnd      = 0.04")
    assert "nOuts     = 4" in src and "tEnd      = 0.04" in src
    outs = {}
    for tag, ng in (("one", 1), ("many", n)):
        d = tmp_path / tag
        d.mkdir()
        (d / "p.txt").write_text(src)
        env = dict(os.environ, PDEODE_SEED="3", PDEODE_NGPUS=str(ng))
        r = subprocess.run([exe], input="p.txt\n", cwd=d, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[tag] = (np.loadtxt(d / "total.dat"), np.array([float(t) for t in open(d / "output.dat").read().split()]),
                     r.stdout)
    ta, tb = outs["one"][0], outs["many"][0]
    assert ta.shape == tb.shape and np.allclose(ta, tb, rtol=1e-9, atol=1e-12)
    fa, fb = outs["one"][1], outs["many"][1]
    assert fa.shape == fb.shape and np.linalg.norm(fa - fb) <= 2e-7 * np.linalg.norm(fa)

    def stats(text):
        return [float(l.split("=")[1]) for l in text.splitlines() if "Iters for iterating" in l]
    assert stats(outs["one"][2]) == stats(outs["many"][2])        # Picard statistics: exact


def test_ensemble_over_several_gpus(tmp_path):
    """pdeode_ensemble with several GPUs: one child process per GPU (each sees only its own
    device), members dealt from one shared counter; every member equals the single run."""
    n = min(gpu_count(), 4)
    if n < 2:
        pytest.skip("2 GPUs needed")
    import __graft_entry__ as ge
    ge.build_host()
    base = open(os.path.join(ROOT, "tests", "data", "square3d.txt")).read()
    files = []
    for k in range(1, 3 * n + 1):
        txt = (base.replace("height    = 0.9", f"height    = {0.9 - 0.03 * k}")
               .replace("nu        = 0.30", f"nu        = {0.30 - 0.005 * k}"))
        (tmp_path / f"{k}.txt").write_text(txt)
        files.append(f"{k}.txt")
    env = dict(os.environ, PDEODE_SEED="1", PDEODE_MEMBERS_PER_GPU="2", PDEODE_ENSEMBLE_TIMING="1",
               PDEODE_DEVICES=",".join(str(i) for i in range(n)))
    exe = os.path.join(ROOT, "pde-odesolver_b200", "bin", "pdeode_ensemble")
    r = subprocess.run([exe] + files, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "one process per GPU" in r.stderr
    assert r.stdout.count("-> ok") == len(files)
    solver = os.path.join(ROOT, "pde-odesolver_b200", "bin", "pdeode_solver")
    for k, f in enumerate(files, start=1):
        single = tmp_path / f"single{k}"
        single.mkdir()
        s = subprocess.run([solver], input=str(tmp_path / f) + "\n", text=True, capture_output=True,
                           cwd=single, env=dict(os.environ, PDEODE_SEED="1"), timeout=300)
        assert s.returncode == 0
        for name in ("total.dat", "COprod.dat", "output.dat"):
            assert open(tmp_path / f"{k}Output" / name).read() == open(single / name).read(), (k, name)
