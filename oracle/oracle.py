"""ctypes binding of oracle/libba_oracle.so (CPU restatement of the reference LM path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libba_oracle.so")


def build(force=False, openmp=True):
    src = os.path.join(HERE, "ba_oracle.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "ba_oracle.h"))):
        return LIB
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", LIB, src, "-lm"]
    if openmp:
        cmd.insert(1, "-fopenmp")
    subprocess.check_call(cmd)
    return LIB


class _Problem(C.Structure):
    _fields_ = [("Pall", C.c_int), ("numP", C.c_int), ("Lall", C.c_int), ("numL", C.c_int),
                ("q", C.c_void_p), ("t", C.c_void_p), ("cam", C.c_void_p), ("Xw", C.c_void_p),
                ("E2", C.c_int), ("idx2", C.c_void_p), ("meas2", C.c_void_p), ("omega2", C.c_void_p),
                ("E3", C.c_int), ("idx3", C.c_void_p), ("meas3", C.c_void_p), ("omega3", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.bao_create.restype = C.c_void_p
        L.bao_create.argtypes = [C.POINTER(_Problem), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.bao_destroy.argtypes = [C.c_void_p]
        L.bao_optimize.restype = C.c_int
        L.bao_optimize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("bao_compute_errors", "bao_max_diagonal"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.bao_build_system.argtypes = [C.c_void_p]
        L.bao_update.argtypes = [C.c_void_p]
        L.bao_solve.restype = C.c_int
        L.bao_solve.argtypes = [C.c_void_p, C.c_double]
        L.bao_compute_scale.restype = C.c_double
        L.bao_compute_scale.argtypes = [C.c_void_p, C.c_double]
        L.bao_chi_sqs.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("bao_nhpl", "bao_nblk", "bao_nmul"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        L.bao_get_hpl_structure.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.bao_get_hsc_structure.argtypes = [C.c_void_p] + [C.c_void_p] * 2
        L.bao_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.bao_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.bao_get_system.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.bao_get_schur.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.bao_get_delta.argtypes = [C.c_void_p] + [C.c_void_p] * 2
        L.bao_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle over a FlatProblem (see cuda-bundle-adjustment_b200/graphio.py)."""

    def __init__(self, prob, rk_type=(0, 0), rk_delta=(0.0, 0.0)):
        self.L = lib()
        self.prob = prob
        self._keep = [np.ascontiguousarray(a) for a in (prob.q, prob.t, prob.cam, prob.Xw, prob.idx2, prob.meas2, prob.omega2,
                                                         prob.idx3, prob.meas3, prob.omega3)]
        k = self._keep
        assert k[4].dtype == np.int32 and k[7].dtype == np.int32
        P = _Problem(prob.Pall, prob.numP, prob.Lall, prob.numL, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]),
                     prob.E2, _p(k[4]), _p(k[5]), _p(k[6]), prob.E3, _p(k[7]), _p(k[8]), _p(k[9]))
        rt = (C.c_int * 2)(*[int(v) for v in rk_type])
        rd = (C.c_double * 2)(*[float(v) for v in rk_delta])
        self.h = self.L.bao_create(C.byref(P), rt, rd)
        self.numP, self.numL, self.Pall, self.Lall, self.E = prob.numP, prob.numL, prob.Pall, prob.Lall, prob.nedges

    def __del__(self):
        if getattr(self, "h", None):
            self.L.bao_destroy(self.h)
            self.h = None

    def optimize(self, niter):
        chi = np.zeros(niter); lam = np.zeros(niter); tr = np.zeros(niter, dtype=np.int32)
        n = self.L.bao_optimize(self.h, niter, _p(chi), _p(lam), _p(tr))
        return chi[:n], lam[:n], tr[:n]

    def compute_errors(self):
        return self.L.bao_compute_errors(self.h)

    def build_system(self):
        self.L.bao_build_system(self.h)

    def max_diagonal(self):
        return self.L.bao_max_diagonal(self.h)

    def solve(self, lam):
        return bool(self.L.bao_solve(self.h, float(lam)))

    def update(self):
        self.L.bao_update(self.h)

    def compute_scale(self, lam):
        return self.L.bao_compute_scale(self.h, float(lam))

    def chi_sqs(self):
        out = np.zeros(self.E)
        self.L.bao_chi_sqs(self.h, _p(out))
        return out

    @property
    def nhpl(self):
        return self.L.bao_nhpl(self.h)

    @property
    def nblk(self):
        return self.L.bao_nblk(self.h)

    @property
    def nmul(self):
        return self.L.bao_nmul(self.h)

    def hpl_structure(self):
        colPtr = np.zeros(self.numL + 1, dtype=np.int32); rowInd = np.zeros(self.nhpl, dtype=np.int32)
        e2h = np.zeros(self.E, dtype=np.int32)
        self.L.bao_get_hpl_structure(self.h, _p(colPtr), _p(rowInd), _p(e2h))
        return colPtr, rowInd, e2h

    def hsc_structure(self):
        rowPtr = np.zeros(self.numP + 1, dtype=np.int32); colInd = np.zeros(self.nblk, dtype=np.int32)
        self.L.bao_get_hsc_structure(self.h, _p(rowPtr), _p(colInd))
        return rowPtr, colInd

    def state(self):
        q = np.zeros((self.Pall, 4)); t = np.zeros((self.Pall, 3)); Xw = np.zeros((self.Lall, 3))
        self.L.bao_get_state(self.h, _p(q), _p(t), _p(Xw))
        return q, t, Xw

    def set_state(self, q, t, Xw):
        q, t, Xw = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, t, Xw))
        self.L.bao_set_state(self.h, _p(q), _p(t), _p(Xw))

    def system(self):
        Hpp = np.zeros((self.numP, 36)); bp = np.zeros((self.numP, 6)); Hll = np.zeros((self.numL, 9))
        bl = np.zeros((self.numL, 3)); Hpl = np.zeros((self.nhpl, 18))
        self.L.bao_get_system(self.h, _p(Hpp), _p(bp), _p(Hll), _p(bl), _p(Hpl))
        return Hpp, bp, Hll, bl, Hpl

    def schur(self):
        Hsc = np.zeros((self.nblk, 36)); bsc = np.zeros((self.numP, 6)); inv = np.zeros((self.numL, 9))
        self.L.bao_get_schur(self.h, _p(Hsc), _p(bsc), _p(inv))
        return Hsc, bsc, inv

    def delta(self):
        xp = np.zeros((self.numP, 6)); xl = np.zeros((self.numL, 3))
        self.L.bao_get_delta(self.h, _p(xp), _p(xl))
        return xp, xl

    def threads(self):
        return self.L.bao_threads()
