"""Python mirror of `class MGPIS` (MGPIS.h:8-38) on top of the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .ddpk import Csr
from .lib import check, load_library

SMOOTH_LEX = 0
SMOOTH_MC = 1
KERNEL_CLASSES = ["spmv", "sweep_fwd", "sweep_bwd", "resid", "restrict", "prolong", "coarse", "vector", "sweep_fwd0"]


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Plan:
    """Host-side stage plan of one level (ddpca_plan_*): no GPU needed."""

    def __init__(self, A: Csr, mode: int = SMOOTH_MC, sub_off=None, tri=False):
        lib = load_library()
        h = C.c_void_p()
        if tri:   # A = strictly lower factor L: wavefronts of I + L (and of I + L^T)
            check(lib.ddpca_plan_create_tri(C.c_int(A.shape[0]), _pi(A.rowptr), _pi(A.colidx), C.byref(h)))
        elif sub_off is None:
            check(lib.ddpca_plan_create(C.c_int(A.shape[0]), _pi(A.rowptr), _pi(A.colidx), C.c_int(mode), C.byref(h)))
        else:   # block-diagonal level: blocks planned in parallel and merged
            so = np.ascontiguousarray(sub_off, dtype=np.int32)
            check(lib.ddpca_plan_create_blocks(C.c_int(A.shape[0]), _pi(A.rowptr), _pi(A.colidx), C.c_int(mode), C.c_int(len(so) - 1), _pi(so), C.byref(h)))
        n, ng, ns = C.c_int(), C.c_int(), C.c_int()
        check(lib.ddpca_plan_sizes(h, C.byref(n), C.byref(ng), C.byref(ns)))
        self.n, self.ngroups, self.nstages = n.value, ng.value, ns.value
        self.perm = np.zeros(self.n, dtype=np.int32)
        self.group_start = np.zeros(self.ngroups + 1, dtype=np.int32)
        self.stage_start = np.zeros(self.nstages + 1, dtype=np.int32)
        check(lib.ddpca_plan_get(h, _pi(self.perm), _pi(self.group_start), _pi(self.stage_start)))
        lib.ddpca_plan_destroy(h)


def hierarchy_pointers(hiers):
    """ctypes argument tuple (n, rowptr, colidx, val, P_rowptr, P_colidx, P_val) for a list of hierarchies
    [(A, P), ...] of equal level count, indexed [s * nlevels + l] as ddpca_mg_create_batch expects.
    The Csr objects must outlive the call that consumes the pointers."""
    nlev = len(hiers[0][0])
    ns = len(hiers)
    if any(len(A) != nlev or len(P) != nlev - 1 for A, P in hiers):
        raise ValueError("hierarchies of one batch must have the same number of levels")
    n = (C.c_int * (ns * nlev))(*[a.shape[0] for A, _ in hiers for a in A])
    ipp = C.POINTER(C.c_int) * (ns * nlev)
    dpp = C.POINTER(C.c_double) * (ns * nlev)
    npr = max(1, ns * (nlev - 1))
    ipp2 = C.POINTER(C.c_int) * npr
    dpp2 = C.POINTER(C.c_double) * npr
    As = [a for A, _ in hiers for a in A]
    Ps = [p for _, P in hiers for p in P]
    return (n, ipp(*[_pi(a.rowptr) for a in As]), ipp(*[_pi(a.colidx) for a in As]), dpp(*[_pd(a.val) for a in As]),
            ipp2(*[_pi(p.rowptr) for p in Ps]), ipp2(*[_pi(p.colidx) for p in Ps]), dpp2(*[_pd(p.val) for p in Ps]))


def setup_dryrun(hiers, smoother: int = SMOOTH_MC):
    """ddpca_mg_setup_dryrun: the host half of the batched hierarchy set-up for [(A, P), ...] without a device.
    Returns {"seconds": {stage: s}, "device_bytes": int, "v2_levels": int, "checksum": int}."""
    nlev = len(hiers[0][0])
    ptrs = hierarchy_pointers(hiers)
    sec = (C.c_double * 9)()
    nbytes, nv2, cs = C.c_long(), C.c_int(), C.c_ulonglong()
    check(load_library().ddpca_mg_setup_dryrun(C.c_int(len(hiers)), C.c_int(nlev), *ptrs, C.c_int(smoother), sec,
                                               C.byref(nbytes), C.byref(nv2), C.byref(cs)))
    names = ["concatenate", "plan", "permute", "layout", "upload", "transfer_permute", "transfer_transpose", "transfer_triples",
             "transfer_upload"]
    return {"seconds": dict(zip(names, list(sec))), "device_bytes": nbytes.value, "v2_levels": nv2.value, "checksum": cs.value}


class MGPIS:
    """Multigrid-preconditioned iterative solver on a B200.

    Same surface as the reference class (MGPIS.h:8-38): fill `maxiLeve`, `consStif`
    (list of Csr, coarsest first) and `realProl`, call `ESTABLISH()`, then
    `CG_SOLV(precSwit, totaForc)` / `MULT_VCYC(level, righHand, resuSolu)`.
    `ESTABLISH` uploads the hierarchy and builds the device layout (it replaces
    the reference's L/D/U split, MGPIS.h:40-53, and the level-0 factorisation the
    reference repeats in every CG_SOLV call, MGPIS.h:185).
    """

    def __init__(self, device: int = 0, smoother: int = SMOOTH_MC):
        self.device = device
        self.smoother = smoother
        self.maxiLeve = -1
        self.consStif: list[Csr] = []
        self.realProl: list[Csr] = []
        self._h = None
        self.last_iterNumb = None
        self.last_resid = None
        self.last_tol = None
        self.nsub = 1
        self._batch = None   # [(A, P), ...] of a batched handle
        self._ntot = None

    @classmethod
    def from_batch(cls, hiers, device=0, smoother=SMOOTH_MC):
        """Several subdomain hierarchies [(A, P), ...] (equal level count) as ONE device hierarchy advanced in
        lock-step (ddpca_mg_create_batch).  CG_SOLV then takes / returns the subdomains' vectors one after the
        other; per-subdomain iteration counts in .last_iters."""
        m = cls(device=device, smoother=smoother)
        m._batch = [(list(A), list(P)) for A, P in hiers]
        m.nsub = len(m._batch)
        m.maxiLeve = len(m._batch[0][0]) - 1
        m.consStif = list(m._batch[0][0])
        m.realProl = list(m._batch[0][1])
        args = hierarchy_pointers(m._batch)
        h = C.c_void_p()
        check(load_library().ddpca_mg_create_batch(C.c_int(device), C.c_int(m.nsub), C.c_int(m.maxiLeve + 1), *args, C.c_int(smoother), C.byref(h)))
        m._h = h
        m._ntot = [sum(A[l].shape[0] for A, _ in m._batch) for l in range(m.maxiLeve + 1)]
        return m

    def batch_result(self):
        """(iters[nsub], resid[nsub], tol_abs[nsub]) of the last CG_SOLV."""
        ns = self.nsub
        it = (C.c_long * ns)()
        rs = (C.c_double * ns)()
        tl = (C.c_double * ns)()
        check(load_library().ddpca_mg_batch_result(self._handle(), None, it, rs, tl))
        return list(it), list(rs), list(tl)

    # -- lifetime ---------------------------------------------------------------------
    def ESTABLISH(self) -> int:
        lib = load_library()
        self.close()
        nlev = self.maxiLeve + 1
        if nlev != len(self.consStif) or len(self.realProl) != max(0, nlev - 1):
            raise ValueError("maxiLeve / consStif / realProl are inconsistent")
        A, P = self.consStif, self.realProl
        n = (C.c_int * nlev)(*[a.shape[0] for a in A])
        ipp = C.POINTER(C.c_int) * nlev
        dpp = C.POINTER(C.c_double) * nlev
        npr = max(1, nlev - 1)
        ipp2 = C.POINTER(C.c_int) * npr
        dpp2 = C.POINTER(C.c_double) * npr
        h = C.c_void_p()
        check(
            lib.ddpca_mg_create(
                C.c_int(self.device), C.c_int(nlev), n,
                ipp(*[_pi(a.rowptr) for a in A]), ipp(*[_pi(a.colidx) for a in A]), dpp(*[_pd(a.val) for a in A]),
                ipp2(*[_pi(p.rowptr) for p in P]), ipp2(*[_pi(p.colidx) for p in P]), dpp2(*[_pd(p.val) for p in P]),
                C.c_int(self.smoother), C.byref(h),
            )
        )
        self._h = h
        return 1

    def close(self):
        if self._h is not None:
            load_library().ddpca_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if self._h is None:
            raise RuntimeError("call ESTABLISH() first")
        return self._h

    def n(self, level=None) -> int:
        level = self.maxiLeve if level is None else level
        if self._ntot is not None:
            return self._ntot[level]
        return self.consStif[level].shape[0]

    # -- solvers (host buffers) ----------------------------------------------------------
    def CG_SOLV(self, precSwit: int, totaForc, rel_tol: float = 1.0e-14, maxit: int | None = None):
        """MGPIS::CG_SOLV (MGPIS.h:163-225). Returns resuSolu; iterNumb in .last_iterNumb."""
        b = _f64(totaForc)
        n = self.n()
        if b.shape[0] != n:
            raise ValueError("totaForc has the wrong length")
        x = np.empty(n)
        it, res, tol = C.c_long(), C.c_double(), C.c_double()
        check(
            load_library().ddpca_mg_pcg(
                self._handle(), C.c_int(precSwit), _pd(b), _pd(x), C.c_double(rel_tol),
                C.c_long((n if self.nsub == 1 else 0) if maxit is None else maxit), C.byref(it), C.byref(res), C.byref(tol),
            )
        )
        self.last_iterNumb, self.last_resid, self.last_tol = it.value, res.value, tol.value
        if self.nsub > 1:
            self.last_iters = self.batch_result()[0]
        return x

    def CG_SOLV_dev(self, precSwit: int, b_ptr: int, x_ptr: int, rel_tol: float = 1.0e-14, maxit: int | None = None):
        """Same with operands already in HBM (raw device pointers, reference numbering)."""
        n = self.n()
        it, res, tol = C.c_long(), C.c_double(), C.c_double()
        check(
            load_library().ddpca_mg_pcg_dev(
                self._handle(), C.c_int(precSwit), C.c_void_p(b_ptr), C.c_void_p(x_ptr), C.c_double(rel_tol),
                C.c_long((n if self.nsub == 1 else 0) if maxit is None else maxit), C.byref(it), C.byref(res), C.byref(tol),
            )
        )
        self.last_iterNumb, self.last_resid, self.last_tol = it.value, res.value, tol.value
        return it.value

    def BiCGSTAB_SOLV(self, precSwit: int, totaForc, rel_tol: float = 1.0e-14, maxit: int | None = None):
        """MGPIS::BiCGSTAB_SOLV (MGPIS.h:350-432). Returns resuSolu; iterNumb in .last_iterNumb."""
        b = _f64(totaForc)
        n = self.n()
        x = np.empty(n)
        it, res, tol = C.c_long(), C.c_double(), C.c_double()
        check(load_library().ddpca_mg_bicgstab(self._handle(), C.c_int(precSwit), _pd(b), _pd(x), C.c_double(rel_tol),
                                               C.c_long(n if maxit is None else maxit), C.byref(it), C.byref(res), C.byref(tol)))
        self.last_iterNumb, self.last_resid, self.last_tol = it.value, res.value, tol.value
        return x

    def GMRES_SOLV(self, precSwit: int, totaForc):
        """MGPIS::GMRES_SOLV (MGPIS.h:227-348): restarted GMRES(10). iterNumb in .last_iterNumb."""
        b = _f64(totaForc)
        x = np.empty(self.n())
        it, res, tol = C.c_long(), C.c_double(), C.c_double()
        check(load_library().ddpca_mg_gmres(self._handle(), C.c_int(precSwit), _pd(b), _pd(x), C.byref(it), C.byref(res), C.byref(tol)))
        self.last_iterNumb, self.last_resid, self.last_tol = it.value, res.value, tol.value
        return x

    def MULT_SOLV(self, totaForc):
        """MGPIS::MULT_SOLV (MGPIS.h:130-160): stand-alone V-cycle iteration."""
        b = _f64(totaForc)
        x = np.empty(self.n())
        it, res = C.c_long(), C.c_double()
        check(load_library().ddpca_mg_mult_solv(self._handle(), _pd(b), _pd(x), C.byref(it), C.byref(res)))
        self.last_iterNumb, self.last_resid = it.value, res.value
        return x

    def MULT_VCYC(self, tempLeve: int, righHand, resuSolu=None):
        """MGPIS::MULT_VCYC (MGPIS.h:55-128); resuSolu is the in/out iterate (zeros if None)."""
        b = _f64(righHand)
        x = np.zeros_like(b) if resuSolu is None else np.array(resuSolu, dtype=np.float64)
        check(load_library().ddpca_mg_vcycle(self._handle(), C.c_int(tempLeve), _pd(b), _pd(x)))
        return x

    # -- single operators (kernel-level parity) --------------------------------------------
    def spmv(self, level: int, x):
        x = _f64(x)
        y = np.empty(self.n(level))
        check(load_library().ddpca_mg_spmv(self._handle(), C.c_int(level), _pd(x), _pd(y)))
        return y

    def restrict(self, level: int, r_fine):
        r = _f64(r_fine)
        out = np.empty(self.n(level))
        check(load_library().ddpca_mg_restrict(self._handle(), C.c_int(level), _pd(r), _pd(out)))
        return out

    def prolong_add(self, level: int, e_coarse, x_fine):
        e = _f64(e_coarse)
        x = np.array(x_fine, dtype=np.float64)
        check(load_library().ddpca_mg_prolong_add(self._handle(), C.c_int(level), _pd(e), _pd(x)))
        return x

    def coarse_solve(self, b):
        b = _f64(b)
        x = np.empty_like(b)
        check(load_library().ddpca_mg_coarse_solve(self._handle(), _pd(b), _pd(x)))
        return x

    # -- measurement ----------------------------------------------------------------------
    def level_info(self, level: int):
        n, nnz, ng, ns = C.c_long(), C.c_long(), C.c_int(), C.c_int()
        check(load_library().ddpca_mg_level_info(self._handle(), C.c_int(level), C.byref(n), C.byref(nnz), C.byref(ng), C.byref(ns)))
        return {"n": n.value, "nnz": nnz.value, "groups": ng.value, "stages": ns.value}

    def launch_count(self, reset: bool = False) -> int:
        return int(load_library().ddpca_mg_launch_count(self._handle(), C.c_int(1 if reset else 0)))

    def set_stream(self, stream_ptr: int):
        check(load_library().ddpca_mg_set_stream(self._handle(), C.c_void_p(stream_ptr)))

    def profile(self, enable: bool):
        check(load_library().ddpca_mg_profile(self._handle(), C.c_int(1 if enable else 0)))

    def profile_get(self):
        """{(kernel_class, level): (ms, launches, algorithmic_bytes)} accumulated since profile(True)."""
        out = {}
        for k, name in enumerate(KERNEL_CLASSES):
            for l in range(self.maxiLeve + 1):
                ms, nl, by = C.c_double(), C.c_long(), C.c_double()
                check(load_library().ddpca_mg_profile_get(self._handle(), C.c_int(k), C.c_int(l), C.byref(ms), C.byref(nl), C.byref(by)))
                if nl.value:
                    out[(name, l)] = (ms.value, nl.value, by.value)
        return out

    def last_timing(self):
        s, a, d = C.c_double(), C.c_double(), C.c_double()
        check(load_library().ddpca_mg_last_timing(self._handle(), C.byref(s), C.byref(a), C.byref(d)))
        return {"solve_ms": s.value, "h2d_ms": a.value, "d2h_ms": d.value}

    # -- convenience -----------------------------------------------------------------------
    @classmethod
    def from_hierarchy(cls, A, P, device=0, smoother=SMOOTH_MC):
        m = cls(device=device, smoother=smoother)
        m.maxiLeve = len(A) - 1
        m.consStif = list(A)
        m.realProl = list(P)
        m.ESTABLISH()
        return m
