"""ctypes loader for the CPU oracle (oracle/liboracle*.so).

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class Params(C.Structure):
    """Layout shared by oracle_params and pdeode_params."""
    _fields_ = [
        ("delta", C.c_double), ("nu", C.c_double), ("kappa", C.c_double),
        ("gama", C.c_double), ("tDel", C.c_double), ("xDel", C.c_double),
        ("e1", C.c_double), ("e2", C.c_double), ("eSoln", C.c_double),
        ("alpha", C.c_int), ("beta", C.c_int), ("dSelect", C.c_int),
        ("fSelect", C.c_int), ("gSelect", C.c_int), ("reserved", C.c_int),
    ]


def default_params(row, **over):
    """parameter.txt:4-29 of the reference, xDel = 1/row (P:110)."""
    p = Params(delta=1e-4, nu=0.30, kappa=0.01, gama=0.60, tDel=1e-3,
               xDel=1.0 / row, e1=1e-8, e2=1e-8, eSoln=1e-8, alpha=4, beta=4,
               dSelect=3, fSelect=1, gSelect=1, reserved=0)
    for k, v in over.items():
        setattr(p, k, v)
    return p


class CgTrace(C.Structure):
    _fields_ = [("rho", C.c_double), ("dot", C.c_double),
                ("alfa", C.c_double), ("err2", C.c_double),
                ("normx", C.c_double)]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR],
                          stdout=subprocess.DEVNULL)


def load(omp=False):
    name = "liboracle_omp.so" if omp else "liboracle.so"
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    lib.oracle_dfunc.restype = C.c_double
    lib.oracle_dfunc.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.oracle_ffunc.restype = C.c_double
    lib.oracle_ffunc.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    lib.oracle_powi.restype = C.c_double
    lib.oracle_powi.argtypes = [C.c_double, C.c_int]
    lib.oracle_gen_matrix.restype = None
    lib.oracle_gen_matrix.argtypes = [_dp, _dp, _dp, _ip, _dp, C.c_int, C.c_int,
                                      C.POINTER(Params)]
    lib.oracle_amuxd.restype = None
    lib.oracle_amuxd.argtypes = [C.c_long, _dp, _dp, _dp, C.c_int, _ip]
    lib.oracle_dotprod.restype = C.c_double
    lib.oracle_dotprod.argtypes = [C.c_long, _dp, _dp]
    lib.oracle_solve_ls_diag.restype = None
    lib.oracle_solve_ls_diag.argtypes = [
        C.c_long, C.c_int, _ip, _dp, _dp, _dp, C.POINTER(C.c_longlong),
        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), _dp,
        C.c_void_p, C.c_long]
    lib.oracle_solve_c.restype = None
    lib.oracle_solve_c.argtypes = [_dp, _dp, _dp, _dp, C.c_long, C.c_double,
                                   C.c_double, C.c_double, C.c_int]
    lib.oracle_calc_diff.restype = C.c_double
    lib.oracle_calc_diff.argtypes = [_dp, _dp, C.c_int, C.c_int]
    lib.oracle_calc_mass.restype = C.c_double
    lib.oracle_calc_mass.argtypes = [_dp, C.c_long]
    lib.oracle_calc_peak_interface.restype = None
    lib.oracle_calc_peak_interface.argtypes = [
        _dp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
        C.POINTER(C.c_double)]
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [C.c_int, C.c_int, C.POINTER(Params)]
    lib.oracle_destroy.restype = None
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_set_state.restype = None
    lib.oracle_set_state.argtypes = [C.c_void_p, _dp, _dp]
    lib.oracle_get_state.restype = None
    lib.oracle_get_state.argtypes = [C.c_void_p, _dp, _dp]
    lib.oracle_set_mnew.restype = None
    lib.oracle_set_mnew.argtypes = [C.c_void_p, _dp]
    lib.oracle_get_mnew.restype = None
    lib.oracle_get_mnew.argtypes = [C.c_void_p, _dp]
    lib.oracle_set_maxit.restype = None
    lib.oracle_set_maxit.argtypes = [C.c_void_p, C.c_longlong]
    lib.oracle_step.restype = C.c_int
    lib.oracle_step.argtypes = [
        C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong),
        C.POINTER(C.c_longlong), C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_first_solve_trace.restype = C.c_longlong
    lib.oracle_first_solve_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_num_threads.argtypes = []
    return lib


class Oracle:
    """Stepper over the locals of solveOrder2 (P:623-628)."""

    def __init__(self, row, col, params, omp=False):
        self.lib = load(omp)
        self.row, self.col, self.n = row, col, row * col
        self.params = params
        self.h = self.lib.oracle_create(row, col, C.byref(params))
        if not self.h:
            raise MemoryError("oracle_create failed")

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_state(self, M, Cc):
        self.lib.oracle_set_state(self.h, np.ascontiguousarray(M, np.float64).ravel(),
                                  np.ascontiguousarray(Cc, np.float64).ravel())

    def get_state(self):
        M = np.empty(self.n); Cc = np.empty(self.n)
        self.lib.oracle_get_state(self.h, M, Cc)
        return M, Cc

    def set_mnew(self, x):
        self.lib.oracle_set_mnew(self.h, np.ascontiguousarray(x, np.float64).ravel())

    def get_mnew(self):
        x = np.empty(self.n)
        self.lib.oracle_get_mnew(self.h, x)
        return x

    def set_maxit(self, maxit):
        self.lib.oracle_set_maxit(self.h, int(maxit))

    def step(self, cap=64):
        cnt = C.c_int(0); s = C.c_longlong(0); mx = C.c_longlong(0)
        nits = (C.c_longlong * cap)()
        diffs = (C.c_double * (2 * cap))()
        rc = self.lib.oracle_step(self.h, C.byref(cnt), C.byref(s), C.byref(mx),
                                  C.cast(nits, C.c_void_p),
                                  C.cast(diffs, C.c_void_p), cap)
        k = min(cnt.value, cap)
        return dict(rc=rc, countIters=cnt.value, nitSum=s.value, nitMax=mx.value,
                    nits=list(nits[:k]),
                    diffs=[(diffs[2 * i], diffs[2 * i + 1]) for i in range(k)])

    def first_solve_trace(self, cap=4096):
        tr = (CgTrace * cap)()
        nit = self.lib.oracle_first_solve_trace(self.h, C.cast(tr, C.c_void_p), cap)
        k = min(nit, cap)
        return nit, [dict(rho=t.rho, dot=t.dot, alfa=t.alfa, err2=t.err2,
                          normx=t.normx) for t in tr[:k]]


# ---- deterministic initial conditions used by the tests and the bench ----

def ic1(row, col, height, depth):
    """MinitialCond = 1 (P:279-288): M = max(0, a*(x*xDel)^4 + height),
    x = (i-1)/col (row index), a = -height/depth^4; C = 1 (P:276)."""
    xDel = 1.0 / row
    a = -height / ((depth * depth) * (depth * depth))     # (depth)**4 as gfortran expands it
    x = np.arange(row, dtype=np.float64)
    xx = x * xDel
    f = a * ((xx * xx) * (xx * xx)) + height
    # (x*xDel)**4 with an integer literal exponent -> (t*t)*(t*t)
    f = np.where(f <= 0, 0.0, f)
    M = np.repeat(f, col)
    return M, np.ones(row * col)


def ic2(row, col, height):
    """MinitialCond = 2 (P:291-292): homogeneous."""
    return np.full(row * col, float(height)), np.ones(row * col)


def ic_blobs(row, col, height, depth, npts, seed):
    """Deterministic stand-in for MinitialCond = 5 (P:328-352) with a fixed
    seed and total = 0: paraboloid inoculation sites on the y=0
    side, normalised so the mean non-zero value is `height`."""
    rng = np.random.RandomState(seed)
    xDel = 1.0 / row
    a = -height / (depth * depth)
    j = np.arange(1, row * col + 1)
    x = (j % col) * xDel
    y = (j // row) * xDel
    M = np.zeros(row * col)
    for _ in range(npts):
        xr, yr = rng.random_sample(), rng.random_sample() * 0.1
        innoc = a * ((x - xr) * (x - xr) + (y - yr) * (y - yr)) + height
        M += np.where(innoc >= 0, innoc, 0.0)
    nz = M > 0
    M = M * height / (M[nz].sum() / nz.sum())
    return M, np.ones(row * col)


# ---- one time step replayed from the building blocks, with every solve's CG trace ----

NEAR = 1.25      # a CG iteration whose err2 came within this factor of the stop threshold
NEAR_CAP = 32


def replay_step(lib, p, row, col, M, Cc, Mnew):
    """One time step (P:700-742) from the oracle's building blocks, with the CG
    trace of every solve.  Returns (nits, near): near[s] lists the iterations of
    solve s where the reference came within NEAR of stopping (CG:90) but did
    not -- where a different summation order may well stop."""
    n = row * col
    Mprev, Cprev = M.copy(), Cc.copy()
    M, Cc, Mnew = M.copy(), Cc.copy(), Mnew.copy()
    diffC = diffM = 1.0
    nits, near = [], []
    cap = 4096
    while diffC + diffM > p.eSoln:
        A = np.zeros(5 * n); rhs = np.zeros(n); ioff = np.zeros(5, np.int32)
        lib.oracle_gen_matrix(Mprev, Cc, A, ioff, rhs, row, col, C.byref(p))
        nit = C.c_longlong(100 * n); e1 = C.c_double(p.e1); e2 = C.c_double(p.e2)
        st = C.c_int(); scr = np.zeros(3 * n); tr = (CgTrace * cap)()
        x0 = Mnew.copy()
        lib.oracle_solve_ls_diag(n, 5, ioff, A, Mnew, rhs, C.byref(nit), C.byref(e1), C.byref(e2),
                                 C.byref(st), scr, C.cast(tr, C.c_void_p), cap)
        k = int(nit.value)
        nits.append(k)
        hits = [i + 1 for i in range(min(k, cap) - 1) if tr[i].err2 < NEAR * p.e2 * tr[i].normx]
        if 0 < k <= cap and tr[k - 1].err2 * NEAR > p.e2 * tr[k - 1].normx:
            # The mirror case: the reference itself stopped with err2 within NEAR of the threshold
            # (from below).  Another summation order may miss that stop; where it can stop next is
            # read off the reference's own CG with the stop rule switched off (e2 = 0, the same
            # start vector): every later iteration within NEAR of the threshold, up to and
            # including the first one clearly below it.
            x2 = x0.copy()
            nit2 = C.c_longlong(min(k + 64, cap)); e1b = C.c_double(p.e1); e2b = C.c_double(0.0)
            tr2 = (CgTrace * cap)()
            lib.oracle_solve_ls_diag(n, 5, ioff, A, x2, rhs, C.byref(nit2), C.byref(e1b), C.byref(e2b),
                                     C.byref(st), scr, C.cast(tr2, C.c_void_p), cap)
            for i in range(k, min(int(nit2.value), cap)):
                thr = p.e2 * tr2[i].normx
                if tr2[i].err2 < NEAR * thr:
                    hits.append(i + 1)
                if tr2[i].err2 * NEAR < thr:
                    break
        near.append(hits if len(hits) <= NEAR_CAP else hits[:NEAR_CAP // 2] + hits[-NEAR_CAP // 2:])
        Cn = np.empty(n)
        lib.oracle_solve_c(Mprev, Mnew, Cprev, Cn, n, p.kappa, p.gama, p.tDel, p.gSelect)
        diffC = lib.oracle_calc_diff(Cc, Cn, row, col)
        diffM = lib.oracle_calc_diff(M, Mnew, row, col)
        Cc = Cn
        M = Mnew.copy()
    return nits, near
