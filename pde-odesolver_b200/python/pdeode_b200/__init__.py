"""ctypes binding of libpdeode_b200.so (include/pdeode_b200.h).

This is plumbing for the tests and bench.py: every call goes straight through
the C ABI.  There is no CPU fallback -- if the CUDA library is missing the
import fails, and without a CUDA device `Solver(...)` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))            # pde-odesolver_b200/
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libpdeode_b200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "pdeode_b200.h")

OK = 0
ERR_CUDA, ERR_NCCL, ERR_ARG, ERR_PICARD_CAP, ERR_NOMEM, ERR_NODEVICE = 1, 2, 3, 4, 5, 6
ERR_MASK = 0xFF
WARN_CG_CAP = 0x100
NCCL_ID_BYTES = 128

ARR_M, ARR_C, ARR_MNEW, ARR_D, ARR_AC, ARR_R, ARR_P, ARR_Q = range(8)
KERNEL_CG_SPMV, KERNEL_CG_UPDATE, KERNEL_INIT, KERNEL_SOLVEC, KERNEL_DIFFCOEF = range(5)
KERNEL_CG_SPMV_X, KERNEL_CG_UPDATE_R = 5, 6


class Params(C.Structure):
    """pdeode_params (P:600-611 of the reference)."""
    _fields_ = [
        ("delta", C.c_double), ("nu", C.c_double), ("kappa", C.c_double),
        ("gama", C.c_double), ("tDel", C.c_double), ("xDel", C.c_double),
        ("e1", C.c_double), ("e2", C.c_double), ("eSoln", C.c_double),
        ("alpha", C.c_int), ("beta", C.c_int), ("dSelect", C.c_int),
        ("fSelect", C.c_int), ("gSelect", C.c_int), ("reserved", C.c_int),
    ]


def default_params(row, **over):
    """parameter.txt:4-29 of the reference; xDel = 1/row (P:110)."""
    p = Params(delta=1e-4, nu=0.30, kappa=0.01, gama=0.60, tDel=1e-3,
               xDel=1.0 / row, e1=1e-8, e2=1e-8, eSoln=1e-8, alpha=4, beta=4,
               dSelect=3, fSelect=1, gSelect=1, reserved=0)
    for k, v in over.items():
        setattr(p, k, v)
    return p


class PdeodeError(RuntimeError):
    def __init__(self, code, what, detail=""):
        self.code = code
        super().__init__(f"{what}: {_lib().pdeode_strerror(code).decode()} [{code}] {detail}")


_LIB = None


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp, llp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_longlong)
    sig = {
        "pdeode_create": [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(Params), C.c_int],
        "pdeode_create_dist": [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(Params), C.c_int,
                               C.c_int, C.c_int, vp],
        "pdeode_create_group": [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(Params), C.c_int, ip],
        "pdeode_nccl_unique_id": [vp],
        "pdeode_destroy": [vp],
        "pdeode_local_rows": [vp, ip, ip],
        "pdeode_slab_rows": [C.c_int, C.c_int, C.c_int, ip, ip],
        "pdeode_device_count": [ip],
        "pdeode_set_stream": [vp, vp],
        "pdeode_set_state": [vp, vp, vp],
        "pdeode_get_state": [vp, vp, vp],
        "pdeode_set_initial_condition": [vp, C.c_int, C.c_double, C.c_double, C.c_int, vp, vp],
        "pdeode_step": [vp, ip, llp, llp],
        "pdeode_step_many": [vp, C.c_int, ip, llp, ip, llp, llp],
        "pdeode_frame_size": [vp, C.c_int, ip, ip],
        "pdeode_get_frame": [vp, C.c_int, vp, vp],
        "pdeode_get_frame_rowmeans": [vp, C.c_int, vp, vp],
        "pdeode_mass": [vp, dp, dp],
        "pdeode_peak": [vp, dp, dp, dp],
        "pdeode_set_maxit": [vp, C.c_longlong],
        "pdeode_set_picard_cap": [vp, C.c_int],
        "pdeode_reconfigure": [vp, C.POINTER(Params)],
        "pdeode_last_step_trace": [vp, llp, dp, C.c_int, ip],
        "pdeode_cg_trace_enable": [vp, C.c_int],
        "pdeode_cg_trace_get": [vp, dp, C.c_int, ip],
        "pdeode_debug_get_array": [vp, C.c_int, vp],
        "pdeode_launch_count": [vp, llp],
        "pdeode_exchange_mode": [vp, ip],
        "pdeode_engine_info": [vp, C.c_char_p, C.c_int],
        "pdeode_time_kernel": [vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
        "pdeode_kernel_bytes_per_cell": [C.c_int],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name in ("pdeode_strerror", "pdeode_last_error", "pdeode_version"):
        getattr(lib, name).restype = C.c_char_p
    lib.pdeode_strerror.argtypes = [C.c_int]
    lib.pdeode_last_error.argtypes = [vp]
    lib.pdeode_version.argtypes = []
    _LIB = lib
    return lib


def exported_symbols():
    """Names declared in include/pdeode_b200.h."""
    import re
    with open(HEADER_PATH) as f:
        txt = f.read()
    return sorted(set(re.findall(r"\b(pdeode_[a-z_0-9]+)\s*\(", txt)))


def nccl_unique_id():
    buf = (C.c_char * NCCL_ID_BYTES)()
    rc = _lib().pdeode_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc:
        raise PdeodeError(rc, "pdeode_nccl_unique_id")
    return bytes(buf)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class Solver:
    """Device-resident stepper: the loop body of solveOrder2 (P:700-742)."""

    def __init__(self, row, col, params, device=0, rank=0, nranks=1, nccl_id=None, ngpus=1,
                 devices=None):
        """ngpus > 1: row slabs over several GPUs of THIS process (pdeode_create_group); the
        object then behaves like a single-GPU solver of the whole grid.  nranks > 1: one slab
        of a multi-process solve (pdeode_create_dist)."""
        self.lib = _lib()
        self.row, self.col = row, col
        self.params = params
        self.h = C.c_void_p()
        if ngpus > 1:
            devs = (C.c_int * ngpus)(*(devices if devices is not None else range(ngpus)))
            rc = self.lib.pdeode_create_group(C.byref(self.h), row, col, C.byref(params), ngpus, devs)
        elif nranks == 1:
            rc = self.lib.pdeode_create(C.byref(self.h), row, col, C.byref(params), device)
        else:
            idbuf = C.create_string_buffer(nccl_id, NCCL_ID_BYTES)
            rc = self.lib.pdeode_create_dist(C.byref(self.h), row, col, C.byref(params), device,
                                             rank, nranks, C.cast(idbuf, C.c_void_p))
        if rc:
            detail = self.lib.pdeode_last_error(self.h).decode() if self.h else ""
            if self.h:
                self.lib.pdeode_destroy(self.h)
                self.h = C.c_void_p()
            raise PdeodeError(rc, "pdeode_create", detail)
        r0, nr = C.c_int(), C.c_int()
        self.lib.pdeode_local_rows(self.h, C.byref(r0), C.byref(nr))
        self.row0, self.nrows = r0.value, nr.value
        self.n_local = self.nrows * col

    def _ck(self, rc, what):
        if rc & ERR_MASK:
            raise PdeodeError(rc, what, self.lib.pdeode_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.pdeode_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        self._ck(self.lib.pdeode_set_stream(self.h, C.c_void_p(stream_ptr)), "set_stream")

    def set_state(self, M, Cc):
        """M, C: this context's slab (nrows*col doubles), host memory."""
        M = np.ascontiguousarray(M, np.float64).reshape(-1)
        Cc = np.ascontiguousarray(Cc, np.float64).reshape(-1)
        assert M.size == self.n_local and Cc.size == self.n_local
        self._ck(self.lib.pdeode_set_state(self.h, _ptr(M), _ptr(Cc)), "set_state")

    def set_state_ptr(self, m_ptr, c_ptr):
        self._ck(self.lib.pdeode_set_state(self.h, C.c_void_p(m_ptr), C.c_void_p(c_ptr)), "set_state")

    def get_state(self, M=None, Cc=None):
        if M is None:
            M = np.empty(self.n_local)
        if Cc is None:
            Cc = np.empty(self.n_local)
        self._ck(self.lib.pdeode_get_state(self.h, _ptr(M), _ptr(Cc)), "get_state")
        return M, Cc

    def get_state_ptr(self, m_ptr, c_ptr):
        self._ck(self.lib.pdeode_get_state(self.h, C.c_void_p(m_ptr), C.c_void_p(c_ptr)), "get_state")

    def set_initial_condition(self, ic, depth, height, px=(), py=()):
        """setInitialConditions on the device (P:262-439)."""
        px = np.ascontiguousarray(px, np.float64); py = np.ascontiguousarray(py, np.float64)
        self._ck(self.lib.pdeode_set_initial_condition(
            self.h, ic, depth, height, px.size, _ptr(px) if px.size else None,
            _ptr(py) if py.size else None), "set_initial_condition")

    def frame(self, filter_):
        """Decimated M, C of this slab, as printToFile writes them (P:834-842)."""
        nr, nc = C.c_int(), C.c_int()
        self._ck(self.lib.pdeode_frame_size(self.h, filter_, C.byref(nr), C.byref(nc)), "frame_size")
        M = np.empty((nr.value, nc.value)); Cc = np.empty((nr.value, nc.value))
        self._ck(self.lib.pdeode_get_frame(self.h, filter_, _ptr(M), _ptr(Cc)), "get_frame")
        return M, Cc

    def frame_rowmeans(self, filter_):
        """Per-row column means, as printToFile2D writes them (P:898-925)."""
        nr, nc = C.c_int(), C.c_int()
        self._ck(self.lib.pdeode_frame_size(self.h, filter_, C.byref(nr), C.byref(nc)), "frame_size")
        m = np.empty(nr.value); c = np.empty(nr.value)
        self._ck(self.lib.pdeode_get_frame_rowmeans(self.h, filter_, _ptr(m), _ptr(c)), "rowmeans")
        return m, c

    def step(self):
        cnt, s, mx = C.c_int(), C.c_longlong(), C.c_longlong()
        rc = self.lib.pdeode_step(self.h, C.byref(cnt), C.byref(s), C.byref(mx))
        if (rc & ERR_MASK) not in (OK, ERR_PICARD_CAP):
            self._ck(rc, "step")
        return dict(rc=rc, countIters=cnt.value, nitSum=s.value, nitMax=mx.value)

    def step_many(self, n):
        """n time steps with one host wait where the engine allows (pdeode_step_many)."""
        done, cmax = C.c_int(), C.c_int()
        csum, nsum, nmax = C.c_longlong(), C.c_longlong(), C.c_longlong()
        rc = self.lib.pdeode_step_many(self.h, n, C.byref(done), C.byref(csum), C.byref(cmax),
                                       C.byref(nsum), C.byref(nmax))
        if (rc & ERR_MASK) not in (OK, ERR_PICARD_CAP):
            self._ck(rc, "step_many")
        return dict(rc=rc, stepsDone=done.value, countSum=csum.value, countMax=cmax.value,
                    nitSum=nsum.value, nitMax=nmax.value)

    def step_trace(self, cap=64):
        info = self.step()
        nits = (C.c_longlong * cap)(); diffs = (C.c_double * (2 * cap))(); cnt = C.c_int()
        self._ck(self.lib.pdeode_last_step_trace(self.h, nits, diffs, cap, C.byref(cnt)), "trace")
        k = min(cnt.value, cap)
        info["nits"] = list(nits[:k])
        info["diffs"] = [(diffs[2 * i], diffs[2 * i + 1]) for i in range(k)]
        return info

    def mass(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.pdeode_mass(self.h, C.byref(a), C.byref(b)), "mass")
        return a.value, b.value

    def peak(self, peak=0.0, height=0.0, intfac=0.0):
        a, b, c = C.c_double(peak), C.c_double(height), C.c_double(intfac)
        self._ck(self.lib.pdeode_peak(self.h, C.byref(a), C.byref(b), C.byref(c)), "peak")
        return a.value, b.value, c.value

    def reconfigure(self, params):
        """New parameters on the same grid (pdeode_reconfigure); set the state afterwards."""
        self._ck(self.lib.pdeode_reconfigure(self.h, C.byref(params)), "reconfigure")
        self.params = params

    def set_picard_cap(self, cap):
        self._ck(self.lib.pdeode_set_picard_cap(self.h, int(cap)), "set_picard_cap")

    def set_maxit(self, maxit):
        self._ck(self.lib.pdeode_set_maxit(self.h, int(maxit)), "set_maxit")

    def cg_trace_enable(self, cap):
        self._ck(self.lib.pdeode_cg_trace_enable(self.h, cap), "cg_trace_enable")

    def cg_trace(self, cap):
        out = (C.c_double * (5 * cap))(); cnt = C.c_int()
        self._ck(self.lib.pdeode_cg_trace_get(self.h, out, cap, C.byref(cnt)), "cg_trace_get")
        k = min(cnt.value, cap)
        keys = ("rho", "dot", "alfa", "err2", "normx")
        return cnt.value, [dict(zip(keys, out[5 * i:5 * i + 5])) for i in range(k)]

    def array(self, which):
        out = np.empty(self.n_local)
        self._ck(self.lib.pdeode_debug_get_array(self.h, which, _ptr(out)), "debug_get_array")
        return out

    def exchange_mode(self):
        v = C.c_int()
        self._ck(self.lib.pdeode_exchange_mode(self.h, C.byref(v)), "exchange_mode")
        return {0: "single", 1: "nccl", 2: "peer"}[v.value]

    def engine_info(self):
        buf = C.create_string_buffer(512)
        self._ck(self.lib.pdeode_engine_info(self.h, buf, 512), "engine_info")
        return dict(kv.split("=", 1) for kv in buf.value.decode().split())

    def launch_count(self):
        v = C.c_longlong()
        self._ck(self.lib.pdeode_launch_count(self.h, C.byref(v)), "launch_count")
        return v.value

    def time_kernel(self, kernel, reps):
        ms = C.c_float()
        self._ck(self.lib.pdeode_time_kernel(self.h, kernel, reps, C.byref(ms)), "time_kernel")
        return ms.value

    @staticmethod
    def bytes_per_cell(kernel):
        return _lib().pdeode_kernel_bytes_per_cell(kernel)
