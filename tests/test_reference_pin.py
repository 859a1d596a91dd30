"""Pins the CPU oracle (oracle/pdeode_oracle.c) and the host program against the
reference ITSELF: oracle/_ref/ holds the reference's three .f90 files translated
statement by statement to C by oracle/f90c/f90toc.py (no Fortran compiler exists
in this image) and compiled with the flags of runAll.sh:6.  Everything here is
bit-exact: same summation order, no FMA, same powi.

CPU only.  Needs oracle/_ref (built by __graft_entry__.build() where
/root/reference exists; it travels to the GPU box with the snapshot); skipped
where neither the built files nor the reference sources exist.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol
import ref_lib as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="oracle/_ref not built and /root/reference absent")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DATA = os.path.join(HERE, "data")


@pytest.fixture(scope="module")
def ref():
    return rl.Ref()


@pytest.fixture(scope="module")
def lib():
    ol.build_oracle()
    return ol.load()


def fields(row, col, seed, lo=0.0, hi=0.95):
    rng = np.random.RandomState(seed)
    return rng.uniform(lo, hi, row * col), rng.uniform(0.05, 1.0, row * col)


# ------------------------------------------------------------------ translator self-checks

def test_translator_language_rules(tmp_path):
    """The translator on a probe that exercises the rules the parity argument rests on:
    integer division, INT truncation, MOD sign, real**integer = powi, serial sum(),
    column-major 2-D indexing, DO bounds evaluated once, copy-in/copy-out dummies."""
    src = tmp_path / "probe.f90"
    src.write_text("""
subroutine probe(n, x, a, out, iout)
    implicit none
    integer, intent(in) :: n
    real, dimension(n), intent(in) :: x
    real, dimension(n,2), intent(inout) :: a
    real, dimension(8), intent(out) :: out
    integer, dimension(6), intent(out) :: iout
    integer :: i, m, e
    real :: s
    iout(1) = 7/2 ; iout(2) = (-7)/2 ; iout(3) = MOD(-7, 3) ; iout(4) = int(-2.7)
    iout(5) = int(2.7) ; e = 3
    out(1) = sum(x*x)
    out(2) = sqrt(sum(x*x))/n
    out(3) = x(2)**e
    out(4) = x(2)**4
    out(5) = 1/real(n)
    out(6) = maxval(abs(x))
    s = 0.
    m = 3
    do i = 1, m
        m = 1        ! the trip count was fixed on entry
        s = s + i
    enddo
    out(7) = s
    a(2,2) = a(1,1) + 0.5
    out(8) = 2.0**(-2)
    iout(6) = i
    if (x(1) .GE. 0 .and. .not. (x(2) < 0)) iout(6) = iout(6) + 100
end subroutine
""")
    out_c = tmp_path / "probe.c"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "f90c", "f90toc.py"), "-o",
                           str(out_c), str(src)])
    so = tmp_path / "probe.so"
    subprocess.check_call(["/usr/bin/gcc", "-O3", "-ffp-contract=off", "-std=gnu11", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "oracle", "f90c"), "-o", str(so), str(out_c),
                           os.path.join(ROOT, "oracle", "f90c", "f90rt.c"), "-lm"])
    p = C.CDLL(str(so))
    n = 5
    x = np.array([0.1, 0.7, 1e-3, 3.0, 0.25])
    a = np.arange(10, dtype=np.float64)          # a(i,j) = a[(i-1) + (j-1)*n]
    out = np.zeros(8); iout = np.zeros(6, np.int32)
    p.probe_(C.byref(C.c_int(n)), x.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
             out.ctypes.data_as(C.c_void_p), iout.ctypes.data_as(C.c_void_p))
    assert list(iout) == [3, -3, -1, -2, 2, 104]
    s = 0.0
    for v in x:
        s = s + v * v
    assert out[0] == s and out[1] == np.sqrt(s) / n
    assert out[2] == ol.load().oracle_powi(0.7, 3)
    assert out[3] == (0.7 * 0.7) * (0.7 * 0.7)
    assert out[4] == 1 / 5.0 and out[5] == 3.0 and out[6] == 6.0 and out[7] == 0.25
    assert a[1 + 1 * n] == a[0] + 0.5


# ------------------------------------------------------------------ subroutine by subroutine

@pytest.mark.parametrize("dSelect", [1, 2, 3])
def test_dfunc_ffunc(ref, lib, dSelect):
    rng = np.random.RandomState(3)
    for M in list(rng.uniform(0, 0.99, 40)) + [0.0, 1e-9, 0.5]:
        for a, b in ((4, 4), (1, 6), (3, 2), (0, 0)):
            assert ref.dfunc(M, 1e-4, a, b, dSelect) == lib.oracle_dfunc(M, 1e-4, a, b, dSelect)
    for f in range(1, 7):
        for M, Cc in zip(rng.uniform(0, 1, 30), rng.uniform(0.01, 1, 30)):
            assert ref.ffunc(M, Cc, 0.01, 0.3, f) == lib.oracle_ffunc(M, Cc, 0.01, 0.3, f)


GRIDS = [(5, 5), (9, 4), (33, 33), (40, 24), (17, 4), (64, 64)]


@pytest.mark.parametrize("row,col", GRIDS)
@pytest.mark.parametrize("dSelect,fSelect", [(3, 1), (2, 3), (1, 5), (3, 6), (3, 2), (2, 4)])
def test_genmatrix_amuxd(ref, lib, row, col, dSelect, fSelect):
    """GenMatrixM (P:517-590) incl. the '>=' quirk cell, and amuxd (P:1044-1105)."""
    p = ol.default_params(row, dSelect=dSelect, fSelect=fSelect, tDel=0.01)
    M, Cc = fields(row, col, 11 * row + col)
    n = row * col
    A_r, ioff_r, rhs_r = ref.gen_matrix(M, Cc, row, col, p)
    A_o = np.zeros(5 * n); ioff_o = np.zeros(5, np.int32); rhs_o = np.zeros(n)
    lib.oracle_gen_matrix(M, Cc, A_o, ioff_o, rhs_o, row, col, C.byref(p))
    assert np.array_equal(ioff_r, ioff_o) and list(ioff_r) == [-col, -1, 0, 1, col]
    assert np.array_equal(A_r, A_o)
    assert np.array_equal(rhs_r, rhs_o)
    # quirk Q1: cell i = n-col (1-based) is treated as a last-row cell
    q = n - col - 1
    assert A_r[4 * n + q] == 0.0 and A_r[0 * n + q] != 0.0
    x = np.random.RandomState(5).standard_normal(n)
    y_r = ref.amuxd(x, A_r, ioff_r)
    y_o = np.zeros(n)
    lib.oracle_amuxd(n, x, y_o, A_o, 5, ioff_o)
    assert np.array_equal(y_r, y_o)
    assert ref.dotprod(x, y_r) == lib.oracle_dotprod(n, x, y_o)


@pytest.mark.parametrize("row,col,height,beta", [(33, 33, 0.9, 4), (40, 24, 0.8, 4), (48, 4, 0.9, 4),
                                                 (25, 25, 0.93, 6)])
def test_solve_ls_diag(ref, lib, row, col, height, beta):
    """solveLSDIAG (CG:1-95): solution, iteration count, err2 and stop code, cold and warm start."""
    p = ol.default_params(row, beta=beta)
    M, Cc = ol.ic1(row, col, height, 0.3)
    n = row * col
    A, ioff, rhs = ref.gen_matrix(M, Cc, row, col, p)
    for x0 in (np.zeros(n), M.copy()):
        sol_r, nit_r, e1_r, e2_r, stop_r = ref.solve_ls_diag(A, ioff, x0, rhs, 100 * n, p.e1, p.e2)
        sol_o = x0.copy()
        nit = C.c_longlong(100 * n); e1 = C.c_double(p.e1); e2 = C.c_double(p.e2); stop = C.c_int(0)
        lib.oracle_solve_ls_diag(n, 5, ioff, A, sol_o, rhs, C.byref(nit), C.byref(e1), C.byref(e2),
                                 C.byref(stop), np.zeros(3 * n), None, 0)
        assert nit_r == nit.value and stop_r == stop.value
        assert e2_r == e2.value and e1_r == e1.value
        assert np.array_equal(sol_r, sol_o)
        assert nit_r >= 2
    # iteration cap (CG:88): maxit = 3 -> 4 iterations, stop code -1
    sol_r, nit_r, _, _, stop_r = ref.solve_ls_diag(A, ioff, np.zeros(n), rhs, 3, p.e1, p.e2)
    sol_o = np.zeros(n)
    nit = C.c_longlong(3); e1 = C.c_double(p.e1); e2 = C.c_double(p.e2); stop = C.c_int(0)
    lib.oracle_solve_ls_diag(n, 5, ioff, A, sol_o, rhs, C.byref(nit), C.byref(e1), C.byref(e2),
                             C.byref(stop), np.zeros(3 * n), None, 0)
    assert (nit_r, stop_r) == (nit.value, stop.value) == (4, -1)
    assert np.array_equal(sol_r, sol_o)


def test_solvec_calcdiff_calcmass_peak(ref, lib):
    row, col = 37, 21
    n = row * col
    M, Cc = fields(row, col, 9)
    Mn, _ = fields(row, col, 10)
    C_r = ref.solve_c(M, Mn, Cc, 0.01, 0.6, 1e-3)
    C_o = np.zeros(n)
    lib.oracle_solve_c(M, Mn, Cc, C_o, n, 0.01, 0.6, 1e-3, 1)
    assert np.array_equal(C_r, C_o)
    # gSelect /= 1 leaves Cnew untouched (P:501)
    assert np.all(ref.solve_c(M, Mn, Cc, 0.01, 0.6, 1e-3, gSelect=2) == 0.0)
    assert ref.calc_diff(Cc, C_r, row, col) == lib.oracle_calc_diff(Cc, C_o, row, col)
    assert ref.calc_mass(M, row, col) == lib.oracle_calc_mass(M, n)
    for field in (M, np.zeros(n), np.where(M > 0.9, 0.93, M * 0.1)):    # ties: '>=' keeps the last cell
        pk = C.c_double(-7.0); h = C.c_double(-7.0); itf = C.c_double(-7.0)
        lib.oracle_calc_peak_interface(field, row, col, C.byref(pk), C.byref(h), C.byref(itf))
        assert ref.calc_peak_interface(field, row, col) == (pk.value, h.value, itf.value)


# ------------------------------------------------------------------ the time loop

def n_steps(tEnd, tDel):
    k = 0
    while k * tDel <= tEnd:        # P:663
        k += 1
    return k


CASES = [
    # name, row, col, true2D, params, IC, tEnd, nOuts
    ("stiff33", 33, 33, 0, dict(), ("ic1", 0.9, 0.3), 0.012, 3),
    ("wave48x4", 48, 4, 1, dict(), ("ic1", 0.9, 0.3), 0.02, 4),
    ("blobs33", 33, 33, 0, dict(tDel=0.05), ("blobs", 0.3, 0.2, 6, 3), 1.0, 5),
    ("rect", 40, 24, 0, dict(fSelect=6, dSelect=2, tDel=0.01), ("ic1", 0.8, 0.4), 0.2, 4),
    ("beta6", 25, 25, 0, dict(beta=6), ("ic1", 0.93, 0.3), 0.004, 2),
]


def case_state(row, col, ic):
    if ic[0] == "ic1":
        return ol.ic1(row, col, ic[1], ic[2])
    return ol.ic_blobs(row, col, ic[1], ic[2], ic[3], ic[4])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_solveorder2_equals_oracle_steps(ref, lib, tmp_path, case):
    """solveOrder2 (P:600-763) run as a whole against oracle_step called once per time step:
    bit-identical fields, identical statistics, identical diagnostics files."""
    name, row, col, true2D, over, ic, tEnd, nOuts = case
    p = ol.default_params(row, **over)
    M0, C0 = case_state(row, col, ic)
    r = ref.solve_order2(M0, C0, row, col, p, tEnd, nOuts, true2D, str(tmp_path))
    steps = n_steps(tEnd, p.tDel)
    o = ol.Oracle(row, col, p)
    o.set_state(M0, C0)
    sumP = sumN = maxP = maxN = 0
    filt = int(tEnd / (nOuts * p.tDel))
    masses = []
    for k in range(steps):
        if k % (filt // 100 + 1) == 0:
            M, Cc = o.get_state()
            masses.append((lib.oracle_calc_mass(M, M.size), lib.oracle_calc_mass(Cc, Cc.size)))
        s = o.step()
        assert s["rc"] == 0
        sumP += s["countIters"]; sumN += s["nitSum"]
        maxP = max(maxP, s["countIters"]); maxN = max(maxN, s["nitMax"])
    M, Cc = o.get_state()
    assert np.array_equal(r["M"], M)
    assert np.array_equal(r["C"], Cc)
    assert r["maxIters"] == maxP and r["maxNit"] == maxN
    assert r["avgNit"] == sumN / sumP and r["avgIters"] == sumP / steps      # P:744-745
    total = np.loadtxt(tmp_path / "total.dat", ndmin=2)
    assert total.shape[0] == len(masses)
    assert np.array_equal(total[:, 1], [m for m, _ in masses])
    assert np.array_equal(total[:, 2], [c for _, c in masses])
    assert os.path.exists(tmp_path / ("2D_output.dat" if true2D else "output.dat"))
    assert os.path.exists(tmp_path / "peakInfo.dat") == bool(true2D)


def numbers(path):
    """All numbers of a text file, in order, plus the positions of its empty lines."""
    vals, blanks = [], []
    with open(path) as f:
        for k, line in enumerate(f):
            if line == "\n":
                blanks.append(k)
            vals += [float(t) for t in line.split()]
    return np.array(vals), blanks


@pytest.mark.parametrize("pfile", ["wave2d.txt", "square3d.txt"])
def test_whole_program_equals_host_program(tmp_path, pfile):
    """`echo FILE | a.out` (runAll.sh:7): the translated reference program against the B200 host
    program driven by the oracle stepper -- every number of every output file, and the frame
    terminators seperateOutput.py:14 keys on."""
    rdir = tmp_path / "ref"; rdir.mkdir()
    mdir = tmp_path / "mine"
    src = os.path.join(DATA, pfile)
    with open(src) as f:
        (rdir / pfile).write_text(f.read())
    env = dict(os.environ, PDEODE_SEED="1")
    out = subprocess.run([rl.program()], input=pfile + "\n", cwd=rdir, env=env, text=True,
                         capture_output=True, check=True).stdout
    subprocess.run([sys.executable, os.path.join(HERE, "run_cpu_program.py"), src, str(mdir)],
                   env=dict(env, ADAPTER_TAG="_pin"), check=True, capture_output=True)
    files = sorted(f for f in os.listdir(rdir) if f.endswith(".dat") and f != "statReport.dat")
    assert files and files == sorted(f for f in os.listdir(mdir) if f.endswith(".dat") and f != "statReport.dat")
    for f in files:
        a, ab = numbers(rdir / f)
        b, bb = numbers(mdir / f)
        assert a.shape == b.shape and np.array_equal(a, b), f
        assert ab == bb, f
    # console: the per-frame lines (P:691-696) and the statistics (P:161-169, wall time aside)
    mine = (mdir / "console.txt").read_text()

    def frame_lines(text):
        lines = text.splitlines()
        h = next(i for i, l in enumerate(lines) if "avgIter" in l)
        e = next(i for i, l in enumerate(lines) if "Statsitcs" in l)
        return [l for l in lines[h + 1:e - 1]]

    assert frame_lines(out) == frame_lines(mine)

    def stats(text):
        return [float(l.split("=")[1]) for l in text.splitlines() if "Iters for" in l]

    assert stats(out) == stats(mine) and len(stats(out)) == 4
    assert stats((rdir / "statReport.dat").read_text()) == stats((mdir / "statReport.dat").read_text())


@pytest.mark.parametrize("ic", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_initial_conditions_equal_host(ref, ic):
    """setInitialConditions (P:262-439) against the host program's generators, with the random
    draws of 4, 5, 7, 10 taken from the same seeded stream."""
    import __graft_entry__ as ge
    from test_host_program import HostParams
    ge.build_host()
    host = C.CDLL(os.path.join(ge.LIBDIR, "libpdeode_hostlogic.so"))
    host.pdeode_host_initial_conditions.argtypes = [C.POINTER(HostParams), C.c_ulonglong, C.c_void_p, C.c_void_p]
    for row, col in ((33, 33), (48, 4), (24, 24)):
        if ic in (3, 4, 5, 9, 10) and row != col:
            continue                     # y = i / row reaches beyond the domain: square grids only
        p = HostParams()
        p.row, p.col, p.n, p.pSize = row, col, row * col, row
        p.MinitialCond, p.numInnocu = ic, 5
        p.depth, p.height, p.xDel = 0.2, 0.3, 1.0 / row
        M = np.empty(p.n); Cc = np.empty(p.n)
        host.pdeode_host_initial_conditions(C.byref(p), 42, M.ctypes.data, Cc.ctypes.data)
        C_r, M_r = ref.set_initial_conditions(row, col, p.depth, p.height, p.xDel, ic, 5, seed=42)
        assert np.array_equal(Cc, C_r)
        assert np.array_equal(M, M_r), (ic, row, col, np.abs(M - M_r).max())
