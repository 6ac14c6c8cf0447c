"""ctypes front-end to the C restatement of the reference (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(ddpca-admm_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.check_call(["make", "-C", _HERE, "-j8", "all"])
    return os.path.join(_HERE, "liboracle.so")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
        _LIB = C.CDLL(path)
        _LIB.orc_mg_create.restype = C.c_void_p
        _LIB.orc_cg_solv.restype = C.c_long
        _LIB.orc_bicgstab.restype = C.c_long
        _LIB.orc_mult_solv.restype = C.c_long
        _LIB.orc_gmres.restype = C.c_long
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleMG:
    """Mirror of class MGPIS (MGPIS.h:8-38) on the CPU restatement."""

    def __init__(self, A, P):
        self.A = A
        self.P = P
        self.nlev = len(A)
        L = lib()
        n = (C.c_int * self.nlev)(*[a.shape[0] for a in A])
        ipp = C.POINTER(C.c_int) * self.nlev
        dpp = C.POINTER(C.c_double) * self.nlev
        np_ = max(1, self.nlev - 1)
        ipp2 = C.POINTER(C.c_int) * np_
        dpp2 = C.POINTER(C.c_double) * np_
        self._keep = (A, P)
        self.h = C.c_void_p(
            L.orc_mg_create(
                self.nlev,
                n,
                ipp(*[_p(a.rowptr, C.c_int) for a in A]),
                ipp(*[_p(a.colidx, C.c_int) for a in A]),
                dpp(*[_p(a.val, C.c_double) for a in A]),
                ipp2(*[_p(p.rowptr, C.c_int) for p in P]) if P else ipp2(),
                ipp2(*[_p(p.colidx, C.c_int) for p in P]) if P else ipp2(),
                dpp2(*[_p(p.val, C.c_double) for p in P]) if P else dpp2(),
            )
        )

    def __del__(self):
        try:
            lib().orc_mg_destroy(self.h)
        except Exception:
            pass

    @property
    def n(self):
        return self.A[-1].shape[0]

    def vcycle(self, level, b, x0=None):
        """MULT_VCYC(level, b, x) -- MGPIS.h:55-128; returns the updated x."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
        lib().orc_vcycle(self.h, C.c_int(level), _p(b, C.c_double), _p(x, C.c_double))
        return x

    def coarse_solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        lib().orc_coarse_solve(self.h, _p(b, C.c_double), _p(x, C.c_double))
        return x

    def cg_solv(self, prec, b):
        """CG_SOLV(precSwit, b, x) -- MGPIS.h:163-225.  Returns (x, iterNumb, resid, tol)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        res = C.c_double()
        tol = C.c_double()
        it = lib().orc_cg_solv(self.h, C.c_long(prec), _p(b, C.c_double), _p(x, C.c_double), C.byref(res), C.byref(tol))
        return x, int(it), res.value, tol.value


def _bicgstab(self, prec, b):
    """BiCGSTAB_SOLV(precSwit, b, x) -- MGPIS.h:350-432.  Returns (x, iterNumb, resid, tol)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    res, tol = C.c_double(), C.c_double()
    it = lib().orc_bicgstab(self.h, C.c_long(prec), _p(b, C.c_double), _p(x, C.c_double), C.byref(res), C.byref(tol))
    return x, int(it), res.value, tol.value


def _mult_solv(self, b):
    """MULT_SOLV(b, x) -- MGPIS.h:130-160.  Returns (x, iterNumb, resid)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    res = C.c_double()
    it = lib().orc_mult_solv(self.h, _p(b, C.c_double), _p(x, C.c_double), C.byref(res))
    return x, int(it), res.value


def _gmres(self, prec, b):
    """GMRES_SOLV(precSwit, b, x) -- MGPIS.h:227-348.  Returns (x, iterNumb, resid, tol)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    res, tol = C.c_double(), C.c_double()
    it = lib().orc_gmres(self.h, C.c_long(prec), _p(b, C.c_double), _p(x, C.c_double), C.byref(res), C.byref(tol))
    return x, int(it), res.value, tol.value


OracleMG.gmres = _gmres
OracleMG.bicgstab = _bicgstab
OracleMG.mult_solv = _mult_solv


def spmv(a, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(a.shape[0])
    lib().orc_spmv(C.c_int(a.shape[0]), _p(a.rowptr, C.c_int), _p(a.colidx, C.c_int), _p(a.val, C.c_double), _p(x, C.c_double), _p(y, C.c_double))
    return y
