"""ctypes loader for oracle/_ref/libref*.so: the reference's own subroutines,
mechanically translated from /root/reference/*.f90 to C by oracle/f90c/f90toc.py
(this image has no Fortran compiler) and compiled with the flags of runAll.sh:6.

Fortran calling convention: every argument by reference, names lower-case + '_'.
Test infrastructure: only tests/ and tests/golden/make_ref_golden.py import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_SRC = "/root/reference"


def available():
    """Built already (it travels to the GPU box), or buildable from /root/reference here."""
    if os.path.exists(os.path.join(REF_DIR, "libref.so")) and os.path.exists(os.path.join(REF_DIR, "ref_serial")):
        return True
    return os.path.exists(os.path.join(REF_SRC, "PDE-ODEsolver.f90"))


def build():
    if os.path.exists(os.path.join(REF_SRC, "PDE-ODEsolver.f90")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def program(omp=False):
    """The reference program: the real gfortran build of runAll.sh:6 where oracle/Makefile could
    make one (a box with a gfortran), else the translated build."""
    build()
    real = os.path.join(REF_DIR, "ref_omp_gfortran" if omp else "ref_serial_gfortran")
    if os.path.exists(real) and os.environ.get("PDEODE_REF_TRANSLATED") != "1":
        return real
    return os.path.join(REF_DIR, "ref_omp" if omp else "ref_serial")


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_I = C.POINTER(C.c_int)
_D = C.POINTER(C.c_double)


def _i(v):
    return C.byref(C.c_int(int(v)))


def _d(v):
    return C.byref(C.c_double(float(v)))


class Ref:
    """The reference's subroutines (PDE-ODEsolver.f90, CGdiag.f90)."""

    def __init__(self, omp=False):
        build()
        self.lib = C.CDLL(os.path.join(REF_DIR, "libref_omp.so" if omp else "libref.so"))
        self.lib.f90_rng_reseed.argtypes = [C.c_ulonglong]

    # P:464-472
    def dfunc(self, M, delta, alpha, beta, dSelect):
        d = C.c_double(0.0)
        self.lib.dfunc_(_d(M), C.byref(d), _d(delta), _i(alpha), _i(beta), _i(dSelect))
        return d.value

    # P:445-458
    def ffunc(self, M, Cc, kappa, nu, fSelect):
        f = C.c_double(0.0)
        self.lib.ffunc_(_d(M), _d(Cc), C.byref(f), _d(kappa), _d(nu), _i(fSelect))
        return f.value

    # P:517-590 -> (MatrixM (n,5) column-major as a flat array, Mioff, Mrhs)
    def gen_matrix(self, M, Cc, row, col, p):
        n = row * col
        A = np.zeros(5 * n); ioff = np.zeros(5, np.int32); rhs = np.zeros(n)
        self.lib.genmatrixm_.argtypes = [_dp, _dp, _dp, _ip, _dp] + [C.c_void_p] * 13
        self.lib.genmatrixm_(np.ascontiguousarray(M), np.ascontiguousarray(Cc), A, ioff, rhs,
                             *[x for x in (
                                 _i(row), _i(col), _i(n), _i(5), _d(p.delta), _d(p.nu), _i(p.alpha),
                                 _i(p.beta), _d(p.kappa), _d(p.tDel), _d(p.xDel), _i(p.dSelect),
                                 _i(p.fSelect))])
        return A, ioff, rhs

    # P:1044-1105
    def amuxd(self, x, diag, ioff):
        n = x.size
        y = np.full(n, np.nan)
        self.lib.amuxd_.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_void_p, _ip]
        self.lib.amuxd_(_i(n), np.ascontiguousarray(x), y,
                        np.ascontiguousarray(diag), _i(ioff.size),
                        np.ascontiguousarray(ioff, np.int32))
        return y

    # CG:99-114
    def dotprod(self, u, v):
        s = C.c_double(0.0)
        self.lib.dotprod_.argtypes = [C.c_void_p, _dp, _dp, C.c_void_p]
        self.lib.dotprod_(_i(u.size), np.ascontiguousarray(u),
                          np.ascontiguousarray(v), C.byref(s))
        return s.value

    # CG:1-95 -> (sol, nit, err1, err2, stopcrit)
    def solve_ls_diag(self, A, ioff, sol0, rhs, maxit, e1, e2):
        n = rhs.size
        sol = np.array(sol0, dtype=np.float64, copy=True)
        nit = C.c_int(int(maxit)); err1 = C.c_double(e1); err2 = C.c_double(e2); stop = C.c_int(0)
        self.lib.solvelsdiag_.argtypes = [C.c_void_p, C.c_void_p, _ip, _dp, _dp, _dp] + [C.c_void_p] * 4
        self.lib.solvelsdiag_(_i(n), _i(5),
                              np.ascontiguousarray(ioff, np.int32), np.ascontiguousarray(A), sol,
                              np.ascontiguousarray(rhs), C.byref(nit),
                              C.byref(err1), C.byref(err2),
                              C.byref(stop))
        return sol, nit.value, err1.value, err2.value, stop.value

    # P:482-511
    def solve_c(self, M, Mnew, Csol, kappa, gama, tDel, gSelect=1):
        n = M.size
        Cnew = np.zeros(n)
        self.lib.solvec_.argtypes = [_dp, _dp, _dp, _dp] + [C.c_void_p] * 5
        self.lib.solvec_(np.ascontiguousarray(M), np.ascontiguousarray(Mnew),
                         np.ascontiguousarray(Csol), Cnew,
                         *[x for x in (_i(n), _d(kappa), _d(gama), _d(tDel), _i(gSelect))])
        return Cnew

    # P:948-963
    def calc_diff(self, a, b, row, col):
        d = C.c_double(0.0)
        self.lib.calcdiff_.argtypes = [C.c_void_p, _dp, _dp, C.c_void_p, C.c_void_p]
        self.lib.calcdiff_(C.byref(d), np.ascontiguousarray(a),
                           np.ascontiguousarray(b), _i(row), _i(col))
        return d.value

    # P:772-791
    def calc_mass(self, X, row, col):
        m = C.c_double(0.0)
        self.lib.calcmass_.argtypes = [_dp] + [C.c_void_p] * 4
        self.lib.calcmass_(np.ascontiguousarray(X), C.byref(m),
                           *[x for x in (_i(X.size), _i(row), _i(col))])
        return m.value

    # P:978-1004 -> (peak, height, intfac); outputs start at the given fill value
    def calc_peak_interface(self, M, row, col, fill=-7.0):
        pk = C.c_double(fill); h = C.c_double(fill); itf = C.c_double(fill)
        self.lib.calcpeakinterface_.argtypes = [_dp] + [C.c_void_p] * 5
        self.lib.calcpeakinterface_(np.ascontiguousarray(M), _i(row),
                                    _i(col), C.byref(pk),
                                    C.byref(h), C.byref(itf))
        return pk.value, h.value, itf.value

    # P:262-439 -> (C, M); random draws restart from `seed` (see f90rt.c)
    def set_initial_conditions(self, row, col, depth, height, xDel, ic, npts, seed=1):
        n = row * col
        Cc = np.zeros(n); M = np.zeros(n)
        self.lib.f90_rng_reseed(seed)
        self.lib.setinitialconditions_.argtypes = [_dp, _dp] + [C.c_void_p] * 8
        self.lib.setinitialconditions_(Cc, M, *[x for x in (
            _i(row), _i(col), _i(n), _d(depth), _d(height), _d(xDel), _i(ic), _i(npts))])
        return Cc, M

    # P:600-763: the whole time loop; runs in `cwd` (it writes total.dat etc. there).
    # -> dict(M, C, avgIters, maxIters, avgNit, maxNit)
    def solve_order2(self, M0, C0, row, col, p, tEnd, nOuts, true2D, cwd):
        n = row * col
        M = np.array(M0, dtype=np.float64, copy=True)
        Cc = np.array(C0, dtype=np.float64, copy=True)
        outs = [C.c_double(0.0) for _ in range(4)]
        e1 = C.c_double(p.e1); e2 = C.c_double(p.e2); nit = C.c_int((100 * n) & 0x7fffffff)
        here = os.getcwd()
        os.chdir(cwd)
        try:
            self.lib.solveorder2_.argtypes = [C.c_void_p] * 6 + [_dp, _dp] + [C.c_void_p] * 20
            self.lib.solveorder2_(
                *[x for x in (_d(tEnd), _i(nOuts), _d(p.tDel), _i(n), _i(row), _i(col))],
                M, Cc,
                *[x for x in (
                    _d(p.xDel), _d(p.gama), _d(p.nu), _i(p.alpha), _i(p.beta), _d(p.kappa), _d(p.delta),
                    _i(5), C.byref(e1), C.byref(e2), C.byref(nit), _d(p.eSoln), _i(p.dSelect),
                    _i(p.fSelect), _i(p.gSelect), C.byref(outs[0]), C.byref(outs[1]),
                    C.byref(outs[2]), C.byref(outs[3]), _i(true2D))])
            self.lib.f90_flush_all()
            self.lib.f90_close_all()
        finally:
            os.chdir(here)
        return dict(M=M, C=Cc, avgIters=outs[0].value, maxIters=outs[1].value,
                    avgNit=outs[2].value, maxNit=outs[3].value)
