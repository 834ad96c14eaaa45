"""ctypes binding of oracle/_ref/libcuba_ref.so = the UNMODIFIED reference compiled for sm_100
(oracle/build_ref.sh).  TEST INFRASTRUCTURE ONLY: GPU parity oracle + bench.py's `--impl reference` arm."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(fp32=False):
    return os.path.join(HERE, "_ref", "libcuba_ref_f32.so" if fp32 else "libcuba_ref.so")


def available(fp32=False):
    return os.path.exists(lib_path(fp32))


class _Problem(C.Structure):
    _fields_ = [("Pall", C.c_int), ("numP", C.c_int), ("Lall", C.c_int), ("numL", C.c_int),
                ("q", C.c_void_p), ("t", C.c_void_p), ("cam", C.c_void_p), ("Xw", C.c_void_p),
                ("E2", C.c_int), ("idx2", C.c_void_p), ("meas2", C.c_void_p), ("omega2", C.c_void_p),
                ("E3", C.c_int), ("idx3", C.c_void_p), ("meas3", C.c_void_p), ("omega3", C.c_void_p)]


_libs = {}


def _lib(fp32):
    if fp32 not in _libs:
        L = C.CDLL(lib_path(fp32))
        L.ref_run.restype = C.c_int
        L.ref_run.argtypes = [C.POINTER(_Problem), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_int] + [C.c_void_p] * 7
        _libs[fp32] = L
    return _libs[fp32]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(prob, niter, rk_type=(0, 0), rk_delta=(0.0, 0.0), warmup=0, fp32=False, want_chisq=False):
    """initialize()+optimize(niter) of the reference on a flat problem.  Returns dict(chi2, q, t, Xw, seconds,
    build_seconds, profile, chisq).  Needs a GPU (returns None when the library reports no device)."""
    L = _lib(fp32)
    keep = [np.ascontiguousarray(a) for a in (prob.q, prob.t, prob.cam, prob.Xw, prob.idx2.astype(np.int32), prob.meas2, prob.omega2,
                                               prob.idx3.astype(np.int32), prob.meas3, prob.omega3)]
    P = _Problem(prob.Pall, prob.numP, prob.Lall, prob.numL, _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]),
                 prob.E2, _p(keep[4]), _p(keep[5]), _p(keep[6]), prob.E3, _p(keep[7]), _p(keep[8]), _p(keep[9]))
    chi = np.zeros(max(niter, 1)); q = np.zeros((prob.Pall, 4)); t = np.zeros((prob.Pall, 3)); Xw = np.zeros((prob.Lall, 3))
    chisq = np.zeros(prob.nedges) if want_chisq else None
    sec = np.zeros(2); prof = np.zeros(8)
    rt = (C.c_int * 2)(*[int(v) for v in rk_type]); rd = (C.c_double * 2)(*[float(v) for v in rk_delta])
    n = L.ref_run(C.byref(P), rt, rd, int(warmup), int(niter), _p(chi), _p(q), _p(t), _p(Xw),
                  None if chisq is None else _p(chisq), _p(sec), _p(prof))
    if n < 0:
        return None
    return dict(chi2=chi[:n], q=q, t=t, Xw=Xw, seconds=float(sec[0]), build_seconds=float(sec[1]), profile=prof, chisq=chisq)
