"""ctypes binding of oracle/c/libbkcpu.so -- the C++/OpenMP CPU baseline of the SH2d PALC path.
Test / measurement infrastructure only (see oracle/__init__.py and oracle/c/bk_cpu_baseline.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
LIB_PATH = os.path.join(_HERE, "libbkcpu.so")


class Opts(C.Structure):
    _fields_ = [("ds", C.c_double), ("dsmin", C.c_double), ("dsmax", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double),
                ("a", C.c_double), ("theta", C.c_double), ("eta", C.c_double), ("max_steps", C.c_int32),
                ("newton_tol", C.c_double), ("newton_maxit", C.c_int32), ("gmres_reltol", C.c_double),
                ("gmres_restart", C.c_int32), ("gmres_maxiter", C.c_int32), ("pc_shift", C.c_double), ("nthreads", C.c_int32)]


def build():
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/c build failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.bkcpu_max_threads.restype = C.c_int32
    return _lib


def calibrated_threads(n):
    """fastest OpenMP thread count for the MGS inner loop on vectors of length n (bkcpu_calibrate_threads)"""
    lib = load()
    lib.bkcpu_calibrate_threads.restype = C.c_int32
    return int(lib.bkcpu_calibrate_threads(C.c_int64(n)))


def make_opts(ds=-1e-3, dsmin=1e-4, dsmax=5e-3, p_min=-1.0, p_max=0.0, a=0.5, theta=0.5, eta=150.0, max_steps=5,
              newton_tol=1e-9, newton_maxit=15, reltol=1e-5, restart=100, maxiter=100, pc_shift=1.0, nthreads=0):
    return Opts(ds, dsmin, dsmax, p_min, p_max, a, theta, eta, max_steps, newton_tol, newton_maxit, reltol, restart, maxiter,
                pc_shift, nthreads)


def newton(dims, lengths, l, nu, u0, tol, maxit, opts):
    lib = load()
    u = np.ascontiguousarray(u0, dtype=np.float64).copy()
    itn, itl = C.c_int32(), C.c_int32()
    st = lib.bkcpu_sh2d_newton(C.c_int32(dims[0]), C.c_int32(dims[1]), C.c_double(lengths[0]), C.c_double(lengths[1]),
                               C.c_double(l), C.c_double(nu), u.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(tol),
                               C.c_int32(maxit), C.byref(opts), C.byref(itn), C.byref(itl))
    if st < 0:
        raise RuntimeError(f"bkcpu_sh2d_newton failed ({st}): grid sizes must be powers of two")
    return u, bool(st), itn.value, itl.value


def palc(dims, lengths, nu, u_start, p_start, opts):
    """-> (rows [list of dicts like oracle.palc.continuation], loop_seconds, step_seconds, u_final, (work_newton, work_linear))"""
    lib = load()
    N = dims[0] * dims[1]
    u = np.ascontiguousarray(u_start, dtype=np.float64)
    rows = np.zeros((opts.max_steps + 2, 5))
    tstep = np.zeros(opts.max_steps + 2)
    ufin = np.zeros(N)
    nrows, wn, wl = C.c_int32(), C.c_int32(), C.c_int32()
    secs = C.c_double()
    dp = C.POINTER(C.c_double)
    st = lib.bkcpu_sh2d_palc(C.c_int32(dims[0]), C.c_int32(dims[1]), C.c_double(lengths[0]), C.c_double(lengths[1]), C.c_double(nu),
                             u.ctypes.data_as(dp), C.c_double(p_start), C.byref(opts), rows.ctypes.data_as(dp), C.byref(nrows),
                             C.byref(secs), tstep.ctypes.data_as(dp), ufin.ctypes.data_as(dp), C.byref(wn), C.byref(wl))
    if st != 0:
        raise RuntimeError(f"bkcpu_sh2d_palc failed ({st})")
    out = [dict(param=float(r[0]), x=float(r[1]), itnewton=int(r[2]), itlinear=int(r[3]), ds=float(r[4]), step=i)
           for i, r in enumerate(rows[: nrows.value])]
    return out, secs.value, tstep[: nrows.value].copy(), ufin, (wn.value, wl.value)


def reset_counters():
    load().bkcpu_reset_counters()


def counters():
    """algorithmic bytes since reset_counters(): (BLAS-1 sweeps, CSR SpMVs)"""
    out = (C.c_double * 2)()
    load().bkcpu_counters(out)
    return float(out[0]), float(out[1])


def triad_gbs(nthreads=0):
    """the host's streaming bandwidth with `nthreads` OpenMP threads (STREAM triad), GB/s"""
    lib = load()
    lib.bkcpu_triad_gbs.restype = C.c_double
    return float(lib.bkcpu_triad_gbs(C.c_int32(nthreads)))
