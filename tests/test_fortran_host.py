"""The reference's own Fortran host with its time step behind the C ABI (INTEGRATION.md route 2).

pde-odesolver_b200/fortran/patch_reference.py rewrites solveOrder2 of the reference's
PDE-ODEsolver.f90 (read where it lies) to call the ISO_C_BINDING module; everything else --
parameter file, initial conditions, cadences, console, output files -- stays the reference's
code.  No Fortran compiler exists here, so the patched source is built through the same
mechanical Fortran-to-C translation as oracle/_ref (tests/fortran_host.py).

CPU: linked against an oracle-backed implementation of the ABI, the patched program must write
the SAME BYTES as the unpatched reference program (oracle/_ref/ref_serial) -- which pins the
call sites of route 2 (what is accumulated where, which state a frame sees, the cadences).
GPU: linked against libpdeode_b200.so; with PDEODE_SUM_ORDER=serial its files equal the
unpatched reference program's byte for byte, and in production order they equal the C++ host
program's (pdeode_solver) to rounding."""
import os
import subprocess

import numpy as np
import pytest

import fortran_host as fh
import ref_lib as rl

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")
CASES = ["wave2d.txt", "square3d.txt"]


def run(exe, pfile, cwd, **env):
    cwd.mkdir(exist_ok=True)
    with open(os.path.join(DATA, pfile)) as f:
        (cwd / pfile).write_text(f.read())
    r = subprocess.run([exe], input=pfile + "\n", cwd=cwd, env=dict(os.environ, PDEODE_SEED="1", **env),
                       text=True, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def dat_files(d):
    return sorted(f for f in os.listdir(d) if f.endswith(".dat") and f != "statReport.dat")


def frame_and_stat_lines(console):
    lines = console.splitlines()
    h = next(i for i, l in enumerate(lines) if "avgIter" in l)
    return [l for l in lines[h + 1:] if "Time to compute" not in l]


def assert_same_bytes(a, b):
    files = dat_files(a)
    assert files and files == dat_files(b)
    for f in files:
        assert (a / f).read_bytes() == (b / f).read_bytes(), f


@pytest.mark.skipif(not (fh.available("cpu") and rl.available()), reason="needs /root/reference or prebuilt files")
@pytest.mark.parametrize("pfile", CASES)
def test_patched_fortran_host_writes_the_reference_bytes(tmp_path, pfile):
    exe = fh.build("cpu")
    mine = run(exe, pfile, tmp_path / "patched")
    ref = run(rl.program(), pfile, tmp_path / "ref")
    assert_same_bytes(tmp_path / "patched", tmp_path / "ref")
    assert frame_and_stat_lines(mine) == frame_and_stat_lines(ref)
    assert len(frame_and_stat_lines(mine)) >= 8


@pytest.mark.skipif(not fh.have_sources(), reason="needs /root/reference")
def test_patch_refuses_a_different_source(tmp_path):
    """An anchor that is missing means another source: nothing is written."""
    import sys
    src = open(os.path.join(fh.REF_SRC, "PDE-ODEsolver.f90")).read().replace("diffC = 1", "diffC = 2")
    (tmp_path / "other.f90").write_text(src)
    r = subprocess.run([sys.executable, os.path.join(fh.FDIR, "patch_reference.py"), str(tmp_path / "other.f90"),
                        "-o", str(tmp_path / "out.f90")], capture_output=True, text=True)
    assert r.returncode != 0 and "not the source" in r.stderr
    assert not (tmp_path / "out.f90").exists()


@pytest.mark.gpu
@pytest.mark.skipif(not (fh.available("b200") and rl.available()), reason="needs /root/reference or prebuilt files")
@pytest.mark.parametrize("pfile", CASES)
def test_fortran_host_on_the_gpu(tmp_path, pfile):
    import __graft_entry__ as ge
    exe = fh.build("b200")
    # (1) serial-order verification mode: the unpatched reference program's bytes
    ser = run(exe, pfile, tmp_path / "serial", PDEODE_SUM_ORDER="serial")
    ref = run(rl.program(), pfile, tmp_path / "ref")
    assert_same_bytes(tmp_path / "serial", tmp_path / "ref")
    assert frame_and_stat_lines(ser) == frame_and_stat_lines(ref)
    # (2) production order: the C++ host program over the same library, host-built fields
    run(exe, pfile, tmp_path / "tree")
    ge.build_host()
    run(os.path.join(ge.BINDIR, "pdeode_solver"), pfile, tmp_path / "cxx", PDEODE_HOST_IC="1")
    files = dat_files(tmp_path / "tree")
    assert files == dat_files(tmp_path / "cxx")
    for f in files:
        a = np.array([float(t) for t in (tmp_path / "tree" / f).read_text().split()])
        b = np.array([float(t) for t in (tmp_path / "cxx" / f).read_text().split()])
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-9, atol=1e-13), f


def test_translator_iso_c_binding_rules(tmp_path):
    """The translator's ISO_C_BINDING rules on a probe: VALUE dummies by value, the others by
    address, type(c_ptr) in and out, a bind(C) type's constructor in component order,
    c_long_long results in real arithmetic, module parameters, iand."""
    import ctypes as C
    import sys
    (tmp_path / "m.f90").write_text("""
module probe_c
    use iso_c_binding
    implicit none
    integer(c_int), parameter :: PROBE_MASK = 255
    type, bind(C) :: pair
        real(c_double) :: a
        integer(c_int) :: k, pad = 0
    end type pair
    interface
        integer(c_int) function probe_open(h, p, scale) bind(C, name="probe_open")
            import :: c_ptr, c_int, c_double, pair
            type(c_ptr), intent(out) :: h
            type(pair), intent(in) :: p
            real(c_double), value :: scale
        end function
        integer(c_int) function probe_run(h, n, x, total, big) bind(C, name="probe_run")
            import :: c_ptr, c_int, c_double, c_long_long
            type(c_ptr), value :: h
            integer(c_int), value :: n
            real(c_double), dimension(*), intent(inout) :: x
            real(c_double), intent(out) :: total
            integer(c_long_long), intent(out) :: big
        end function
    end interface
end module probe_c
""")
    (tmp_path / "s.f90").write_text("""
subroutine drive(n, x, out)
    use probe_c
    implicit none
    integer, intent(in) :: n
    real, dimension(n), intent(inout) :: x
    real, dimension(4), intent(out) :: out
    type(c_ptr) :: h
    type(pair) :: p
    integer(c_int) :: rc
    integer(c_long_long) :: big
    real :: total, acc
    p = pair(1.5, n, 0)
    rc = probe_open(h, p, 2.0)
    out(1) = rc
    rc = probe_run(h, n, x, total, big)
    out(2) = total
    acc = 0.5
    acc = acc + big
    out(3) = acc
    out(4) = iand(rc, PROBE_MASK)
end subroutine
""")
    (tmp_path / "c.c").write_text("""
#include <stdlib.h>
struct pair { double a; int k, pad; };
struct st { double a, scale; int k; };
int probe_open(void **h, const struct pair *p, double scale) {
    struct st *s = malloc(sizeof *s); s->a = p->a; s->k = p->k; s->scale = scale; *h = s; return 7; }
int probe_run(void *h, int n, double *x, double *total, long long *big) {
    struct st *s = h; double t = 0; for (int i = 0; i < n; ++i) { x[i] = x[i] * s->scale + s->a; t = t + x[i]; }
    *total = t; *big = 5000000000LL + s->k; free(s); return 0x1234; }
""")
    gen = tmp_path / "gen.c"
    f90c = os.path.join(fh.ROOT, "oracle", "f90c")
    subprocess.check_call([sys.executable, os.path.join(f90c, "f90toc.py"), "-o", str(gen),
                           str(tmp_path / "m.f90"), str(tmp_path / "s.f90")])
    so = tmp_path / "probe.so"
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-I" + f90c, "-o", str(so), str(gen),
                           str(tmp_path / "c.c"), os.path.join(f90c, "f90rt.c"), "-lm"])
    lib = C.CDLL(str(so))
    x = np.array([1.0, 2.0, 3.0])
    out = np.zeros(4)
    lib.drive_(C.byref(C.c_int(3)), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert list(x) == [3.5, 5.5, 7.5]
    assert list(out) == [7.0, 16.5, 5000000003.5, float(0x34)]
