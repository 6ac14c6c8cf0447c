"""Python mirror of the reference's `class MCONTACT` ADMM driver (MCONTACT.h:8-95, :2493-2845)
on top of the C ABI (ddpca_admm_* / ddpca_ldlt_* in include/ddpca_b200.h).

The per-iteration work runs on the device; this file keeps what the reference keeps on the
host around it: the loop of CONTACT_ANALYSIS, MONITOR's ring buffers / VECT_MEDI_OSCI /
MULT_MAXI logic, and resuMoni rows.  Operators come from the reference's own host setup
(MCONTACT::ESTABLISH) -- here through a DDPK dump written by a reference-built driver."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ddpk
from .lib import DdpcaError, check, load_library
from .mgpis import MGPIS, SMOOTH_MC, _pd, _pi, hierarchy_pointers

OPS = ["systTran", "systTran_pena", "inteMass", "inteMass_pena", "inpoLagr", "inteInpo", "pemaInpo_r",
       "globTran", "globTran_pena", "globTran_D", "globTran_1"]
SOLVER_MASS, SOLVER_MASS_PENA = 0, 1


def VECT_MEDI_OSCI(v):
    """PREP.h:147-153."""
    mx, mn = max(v), min(v)
    return (mx + mn) / 2.0, mx - mn


class DIRE_SOLV:
    """A factorised Eigen::SimplicialLDLT (typedef DIRE_SOLV, PREP.h:107) whose solve phase runs
    on the device: perm = permutationP().indices(), L = strictly-lower unit factor (CSR), D = vectorD()."""

    # up to this size an SPD operator is inverted densely on the device (8 n^2 bytes; all interface mass solves of a rank
    # are then ONE block-diagonal product per update); beyond: staged sparse triangular solves with the host factor
    DENSE_MAX = 8192
    # The coarse problems (macroscopic / interface-eliminated) are solved once per ADMM iteration on EVERY rank.  One dense
    # product is faster per solve than the staged sparse sweeps (0.9 ms against ~3 ms at 28 k rows), but the inversion
    # costs 2 n^3 flops at set-up (5 s at 28 k rows against ~1.5 s for uploading the host factor): measured on BLOCK
    # globLeve 3 the dense form only pays for itself below ~12 k rows or over many solves.
    DENSE_MAX_COARSE = 12288

    @classmethod
    def dense(cls, A: ddpk.Csr, device: int = 0):
        """Small SPD operator given as a matrix: dense inverse on the device, no host factor needed."""
        self = cls.__new__(cls)
        self.n = int(A.shape[0])
        h = C.c_void_p()
        check(load_library().ddpca_ldlt_create_dense(C.c_int(device), C.c_int(self.n), _pi(A.rowptr), _pi(A.colidx), _pd(A.val), C.byref(h)))
        self._h = h
        self._owned = True
        return self

    def __init__(self, perm, L: ddpk.Csr, D, device: int = 0):
        lib = load_library()
        self.n = int(L.shape[0])
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        D = np.ascontiguousarray(D, dtype=np.float64)
        h = C.c_void_p()
        check(lib.ddpca_ldlt_create(C.c_int(device), C.c_int(self.n), _pi(perm), _pi(L.rowptr), _pi(L.colidx), _pd(L.val), _pd(D), C.byref(h)))
        self._h = h
        self._owned = True

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        check(load_library().ddpca_ldlt_solve(self._h, _pd(b), _pd(x)))
        return x

    def info(self):
        n, nnz, sf, sb = C.c_int(), C.c_long(), C.c_int(), C.c_int()
        check(load_library().ddpca_ldlt_info(self._h, C.byref(n), C.byref(nnz), C.byref(sf), C.byref(sb)))
        return {"n": n.value, "nnzL": nnz.value, "stages_fwd": sf.value, "stages_bwd": sb.value}

    def release(self):
        """Hand the device object over to an ADMM handle (which then owns it)."""
        self._owned = False
        return self._h

    def close(self):
        if self._owned and self._h is not None:
            load_library().ddpca_ldlt_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _factor_from_dump(d, name, device, fallback_matrix=None, factorize=None, dense_max=None):
    dense_max = DIRE_SOLV.DENSE_MAX if dense_max is None else dense_max
    if fallback_matrix is not None and fallback_matrix.shape[0] <= dense_max:
        return DIRE_SOLV.dense(fallback_matrix, device)
    if name + ".perm" in d:
        return DIRE_SOLV(d[name + ".perm"], ddpk.get_csr(d, name + ".L"), d[name + ".D"], device)
    if factorize is None:
        raise DdpcaError(f"{name}: no factorisation in the dump and no host factoriser given")
    perm, L, D = factorize(fallback_matrix)
    return DIRE_SOLV(perm, L, D, device)


def gamma_project(t, gapTerm, fricCoef: float, device: int = 0):
    """The loop's projection kernel on host arrays (kernel-level parity entry, ddpca_gamma_project):
    inpoGamm = proj(0.5 (t - gapTerm)) and fricStat, MCONTACT.h:2636-2668."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    gap = np.ascontiguousarray(gapTerm, dtype=np.float64)
    d = 1 if fricCoef == 0.0 else 3
    nip = t.shape[0] // d
    g = np.empty_like(t)
    st = np.empty(t.shape[0], dtype=np.int32)
    check(load_library().ddpca_gamma_project(C.c_int(device), C.c_int(nip), C.c_int(d), C.c_double(fricCoef), _pd(t), _pd(gap), _pd(g),
                                             st.ctypes.data_as(C.POINTER(C.c_int))))
    return g, st


class MCONTACT:
    """Multibody contact / domain decomposition ADMM solver on a B200.

    Public state follows the reference: `resuDisp[v]`, `inteAuxi[ts][tv]`, `inteLagr[ts][tv]`,
    `iterNumbReco`, `muscSett`, `fricCoef`, `contBody` (MCONTACT.h:11-58)."""

    def __init__(self, device: int = 0, smoother: int = SMOOTH_MC):
        self.device = device
        self.smoother = smoother
        self._h = None
        self.muscSett = 0
        self.contBody = []
        self.fricCoef = []
        self.MULT_MAXI = 1000  # PREP.h:75
        self.iterNumbReco = None
        self.resuMoni = []
        self.cg_iters = 0
        self.cg_dof_iters = 0.0
        self.body_dof = []
        self.nfull = []

    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_ddpk(cls, d: dict, device: int = 0, smoother: int = SMOOTH_MC, muscSett=None, factorize=None,
                  body_rank=None, rank: int = 0, comm=None, macro_mgpis=None, iterative_above=None, macro1_mgpis=None):
        """Upload everything MCONTACT::ESTABLISH built (dumped by oracle/ref_drivers/admm_hook.h).

        Multi-GPU (one process per GPU): `body_rank[v]` = owning rank (see partition.py), `rank` = this
        process, `comm` = an object with `allreduce_sum(torch_tensor)` and `torch` device tensors for the
        three exchange buffers (see comm.py); only the bodies / sides of this rank are uploaded.

        `macro_mgpis`: an established MGPIS whose finest level is globCoup -- the macroscopic problem is then
        solved by mgpi.CG_SOLV(1, .) like the reference does beyond DIRE_MAXI rows (MCONTACT.h:2560-2562)
        instead of the factor coarSolv_D; ownership of its device hierarchy moves to this object.

        `iterative_above`: interface sides whose mass matrices have more rows than this get no factor; their mass
        systems are solved by batched Jacobi-PCG on the device (the reference: Eigen CG from DIRE_MAXI = 120 000 rows,
        MCONTACT.h:2678-2683).  Default: only when the dump carries no factor for the side."""
        import time

        lib = load_library()
        self = cls(device, smoother)
        tm = {"bodies": 0.0, "side_operators": 0.0, "side_solvers": 0.0, "coarse_solvers": 0.0, "finalize": 0.0}
        t_ = time.time()
        nb, ni = int(d["nbody"][0]), int(d["niface"][0])
        self.nb, self.ni = nb, ni
        self.muscSett = int(d["muscSett"][0]) if muscSett is None else muscSett
        self.rank = rank
        self.comm = comm
        self.body_rank = [0] * nb if body_rank is None else [int(r) for r in body_rank]
        if body_rank is None:
            self.rank = rank = 0
        h = C.c_void_p()
        check(lib.ddpca_admm_create(C.c_int(device), C.c_int(nb), C.c_int(ni), C.c_int(self.muscSett), C.byref(h)))
        self._h = h
        check(lib.ddpca_admm_set_smoother(h, C.c_int(smoother)))
        keep = []   # hierarchy arrays must stay alive until ddpca_admm_finalize
        if body_rank is not None:
            br = (C.c_int * nb)(*self.body_rank)
            check(lib.ddpca_admm_set_partition(h, br, C.c_int(rank)))
        for v in range(nb):
            p = f"body{v}."
            self.body_dof.append(int(d[p + f"consStif{int(d[p + 'maxiLeve'][0])}.shape"][0]))
            self.nfull.append(int(d[p + "nfull"][0]))
            if self.body_rank[v] != rank:
                continue
            L = int(d[p + "maxiLeve"][0])
            A = [ddpk.get_csr(d, p + f"consStif{l}") for l in range(L + 1)]
            P = [ddpk.get_csr(d, p + f"realProl{l}") for l in range(L)]
            args = hierarchy_pointers([(A, P)])
            keep.append((A, P, args))
            F = ddpk.get_csr(d, p + "forcOper")
            nfull = int(d[p + "nfull"][0])
            consForc = np.ascontiguousarray(d[p + "consForc"])
            dispCons = np.ascontiguousarray(d[p + "dispCons"])
            check(lib.ddpca_admm_set_body(h, C.c_int(v), C.c_int(L + 1), *args, C.c_int(nfull), _pd(consForc), _pi(F.rowptr), _pi(F.colidx), _pd(F.val), _pd(dispCons)))
            if self.muscSett & 3:
                a = ddpk.get_csr(d, p + "accuProl")
                check(lib.ddpca_admm_set_body_accuprol(h, C.c_int(v), C.c_int(a.shape[0]), C.c_int(a.shape[1]), _pi(a.rowptr), _pi(a.colidx), _pd(a.val)))
            if self.muscSett & 2:   # interface-eliminated coarse problem, MCONTACT.h:2583
                a = ddpk.get_csr(d, p + "globTran_D_1")
                check(lib.ddpca_admm_set_body_globtran_d1(h, C.c_int(v), C.c_int(a.shape[0]), C.c_int(a.shape[1]), _pi(a.rowptr), _pi(a.colidx), _pd(a.val)))
        tm["bodies"] = time.time() - t_
        self.nc = []
        self.ng = []
        for ts in range(ni):
            p = f"if{ts}."
            cb = [int(x) for x in d[p + "contBody"]]
            fric = float(d[p + "fricCoef"][0])
            nip = int(d[p + "nip"][0])
            gap = np.ascontiguousarray(d[p + "gapTerm"])
            self.contBody.append(cb)
            self.fricCoef.append(fric)
            self.ng.append(gap.shape[0])
            check(lib.ddpca_admm_set_interface(h, C.c_int(ts), C.c_int(cb[0]), C.c_int(cb[1]), C.c_double(fric), C.c_int(nip), _pd(gap)))
            ncs = []
            for tv in range(2):
                q = p + f"s{tv}."
                ncs.append(int(d[q + "inteMass.shape"][0]))
                if self.body_rank[cb[tv]] != rank:
                    continue
                ops = list(range(7)) + ([7, 8, 9] if (self.muscSett & 1) else []) + ([10] if (self.muscSett & 2) else [])
                t_ = time.time()
                for k in ops:
                    m = ddpk.get_csr(d, q + OPS[k])
                    check(lib.ddpca_admm_set_side_op(h, C.c_int(ts), C.c_int(tv), C.c_int(k), C.c_int(m.shape[0]), C.c_int(m.shape[1]), _pi(m.rowptr), _pi(m.colidx), _pd(m.val)))
                tm["side_operators"] += time.time() - t_
                t_ = time.time()
                no_factor = (q + "inteDiso.perm") not in d and ncs[-1] > DIRE_SOLV.DENSE_MAX and factorize is None
                if no_factor or (iterative_above is not None and ncs[-1] > iterative_above):
                    check(lib.ddpca_admm_set_side_iterative(h, C.c_int(ts), C.c_int(tv)))
                    tm["side_solvers"] += time.time() - t_
                    continue
                for which, nm, mat in ((SOLVER_MASS, "inteDiso", "inteMass"), (SOLVER_MASS_PENA, "inteDiso_pena", "inteMass_pena")):
                    s = _factor_from_dump(d, q + nm, device, ddpk.get_csr(d, q + mat), factorize)
                    check(lib.ddpca_admm_set_side_solver(h, C.c_int(ts), C.c_int(tv), C.c_int(which), s.release()))
                tm["side_solvers"] += time.time() - t_
            self.nc.append(ncs)
        t_ = time.time()
        if self.muscSett & 1:
            base = np.ascontiguousarray(d["baseReco"], dtype=np.int64)
            if macro_mgpis is not None:
                nglob = int(ddpk.get_csr(d, "globCoup").shape[0])
                check(lib.ddpca_admm_set_macro_mg(h, C.c_int(nglob), base.ctypes.data_as(C.POINTER(C.c_long)), macro_mgpis._h))
                macro_mgpis._h = None  # ownership moved to the ADMM handle
            else:
                s = _factor_from_dump(d, "coarSolv_D", device, ddpk.get_csr(d, "globCoup"), factorize, DIRE_SOLV.DENSE_MAX_COARSE)
                check(lib.ddpca_admm_set_macro(h, C.c_int(s.n), base.ctypes.data_as(C.POINTER(C.c_long)), s.release()))
        if self.muscSett & 2:   # MCONTACT::MULTISCALE_1 (MCONTACT.h:1672-2343), applied at :2575-2607
            base = np.ascontiguousarray(d["baseReco"], dtype=np.int64)
            gf1 = np.ascontiguousarray(d["globForc_1"], dtype=np.float64)
            if macro1_mgpis is not None:   # MCONTACT.h:2593-2595: mgpi_1.CG_SOLV(1, .); ownership moves to the ADMM handle
                n1 = int(ddpk.get_csr(d, "globCoup_1").shape[0])
                check(lib.ddpca_admm_set_macro1_mg(h, C.c_int(n1), base.ctypes.data_as(C.POINTER(C.c_long)), _pd(gf1), macro1_mgpis._h))
                macro1_mgpis._h = None
            else:
                s1 = _factor_from_dump(d, "coarSolv_D_1", device, ddpk.get_csr(d, "globCoup_1"), factorize, DIRE_SOLV.DENSE_MAX_COARSE)
                check(lib.ddpca_admm_set_macro1(h, C.c_int(s1.n), base.ctypes.data_as(C.POINTER(C.c_long)), _pd(gf1), s1.release()))
        tm["coarse_solvers"] = time.time() - t_
        t_ = time.time()
        if comm is not None:
            ng_, nt_, nm_ = C.c_long(), C.c_long(), C.c_long()
            check(lib.ddpca_admm_exchange_sizes(h, C.byref(ng_), C.byref(nt_), C.byref(nm_)))
            self._xbuf = comm.alloc(ng_.value, nt_.value, nm_.value)   # (globForc, trace_send, trace_recv, moni) device tensors
            ptr = [C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None for t in self._xbuf]
            check(lib.ddpca_admm_set_exchange(h, ptr[0], ptr[1], ptr[2], ptr[3]))
            check(lib.ddpca_admm_set_stream(h, C.c_void_p(comm.stream_ptr())))
        check(lib.ddpca_admm_finalize(h))
        tm["finalize"] = time.time() - t_
        self.upload_times = {k: round(v, 3) for k, v in tm.items()}
        del keep
        self._peers = []
        if comm is not None:
            np_ = C.c_int()
            check(lib.ddpca_admm_exchange_peers(h, C.byref(np_), None, None, None))
            k = np_.value
            pr, off, cnt = (C.c_int * max(1, k))(), (C.c_long * max(1, k))(), (C.c_long * max(1, k))()
            check(lib.ddpca_admm_exchange_peers(h, C.byref(np_), pr, off, cnt))
            self._peers = [(int(pr[i]), int(off[i]), int(cnt[i])) for i in range(k)]
        self.row_len = int(lib.ddpca_admm_row_length(h))
        self.moniReco = [[0.0] * 10 for _ in range(nb + 4 * ni)]  # MCONTACT.h:2494-2498
        return self

    def close(self):
        if self._h is not None:
            load_library().ddpca_admm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def step(self, tc: int):
        """One pass of the loop body (MCONTACT.h:2505-2704) on the device; returns the monitor row."""
        row = np.empty(self.row_len)
        it, dofit = C.c_long(), C.c_double()
        macro = 1 if ((self.muscSett & 3) and tc <= self.MULT_MAXI) else 0  # :2540, :2575
        lib = load_library()
        if self.comm is None:
            check(lib.ddpca_admm_step(self._h, C.c_int(macro), _pd(row), C.byref(it), C.byref(dofit)))
        else:
            # phases of the loop body with the three exchanges of SURVEY.md §8e in between
            gl, tx, rx, mo = self._xbuf
            check(lib.ddpca_admm_phase(self._h, C.c_int(0)))            # body solves (local bodies)
            if macro and (self.muscSett & 1):
                check(lib.ddpca_admm_phase(self._h, C.c_int(1)))        # partial coarse right-hand side
                self.comm.allreduce_sum(gl)
                check(lib.ddpca_admm_phase(self._h, C.c_int(2)))        # replicated coarse solve + correction
            if macro and (self.muscSett & 2):
                check(lib.ddpca_admm_phase(self._h, C.c_int(6)))        # the same for the interface-eliminated problem (:2575-2607)
                self.comm.allreduce_sum(gl)
                check(lib.ddpca_admm_phase(self._h, C.c_int(7)))
            check(lib.ddpca_admm_phase(self._h, C.c_int(3)))            # interface side traces
            if self._peers:
                self.comm.swap(tx, rx, self._peers)                     # pairwise with the owners of the other sides
            check(lib.ddpca_admm_phase(self._h, C.c_int(4)))            # projection, auxiliary and multiplier updates
            check(lib.ddpca_admm_phase(self._h, C.c_int(5)))            # MONITOR sums
            self.comm.allreduce_sum(mo)
            check(lib.ddpca_admm_monitor_row(self._h, _pd(row), C.byref(it), C.byref(dofit)))
        self.cg_iters += it.value
        self.cg_dof_iters += dofit.value
        return row

    def MONITOR(self, tc: int, row):
        """MCONTACT::MONITOR (MCONTACT.h:2725-2845) on the sums the device returned.  Returns 1 when
        converged, -1 otherwise; lowers MULT_MAXI once the oscillation test passes (:2838-2840)."""
        cyc = 10
        flag0 = tc >= cyc
        flag1 = True
        c = 0
        for v in range(self.nb):
            dv, al = row[c], row[c + 1]
            c += 2
            self.moniReco[v][tc % cyc] = dv
            if tc >= cyc:
                medi, osci = VECT_MEDI_OSCI(self.moniReco[v])
                if osci > 0.1 * medi:
                    flag0 = False
            if dv > 1.0e-12 * al:
                flag1 = False
        for ts in range(self.ni):
            for tv in range(2):
                k = self.nb + 4 * ts + 2 * tv
                da, aa, dl, la = row[c], row[c + 1], row[c + 2], row[c + 3]
                c += 4
                self.moniReco[k][tc % cyc] = da
                if tc >= cyc:
                    medi, osci = VECT_MEDI_OSCI(self.moniReco[k])
                    if osci > 0.1 * medi:
                        flag0 = False
                if da > 1.0e-12 * aa:
                    flag1 = False
                self.moniReco[k + 1][tc % cyc] = dl  # multiplier criteria are disabled in the reference (:2822,:2830)
        if flag0:
            self.MULT_MAXI = tc
        return 1 if flag1 else -1

    def CONTACT_ANALYSIS(self, maxiIter: int = 3000):
        """MCONTACT::CONTACT_ANALYSIS (MCONTACT.h:2493-2723).  Returns 1; iterNumbReco as the reference."""
        self.resuMoni = []
        tc = 0
        while tc < maxiIter:
            row = self.step(tc)
            self.resuMoni.append(row)
            if self.MONITOR(tc, row) == 1:
                break
            tc += 1
        self.iterNumbReco = tc
        return 1

    # ---- state read-back (the reference's public members) -----------------------------------------
    @property
    def resuDisp(self):
        out = []
        for v in range(self.nb):
            if self.body_rank[v] != self.rank:
                out.append(None)   # lives on another rank
                continue
            a = np.empty(self.nfull[v])
            check(load_library().ddpca_admm_get_disp(self._h, C.c_int(v), _pd(a)))
            out.append(a)
        return out

    def _side(self, which):
        out = []
        for ts in range(self.ni):
            row = []
            for tv in range(2):
                if self.body_rank[self.contBody[ts][tv]] != self.rank:
                    row.append(None)
                    continue
                a = np.empty(self.nc[ts][tv])
                args = (_pd(a), None) if which == 0 else (None, _pd(a))
                check(load_library().ddpca_admm_get_side(self._h, C.c_int(ts), C.c_int(tv), *args))
                row.append(a)
            out.append(row)
        return out

    @property
    def inteAuxi(self):
        return self._side(0)

    @property
    def inteLagr(self):
        return self._side(1)

    def inpoGamm(self, ts: int):
        """(inpoGamm[ts], fricStat) of the last iteration: the content of resuCont_<ts>.txt."""
        g = np.empty(self.ng[ts])
        st = np.empty(self.ng[ts], dtype=np.int32)
        check(load_library().ddpca_admm_get_gamma(self._h, C.c_int(ts), _pd(g), st.ctypes.data_as(C.POINTER(C.c_int))))
        return g, st

    def launch_count(self, reset=False):
        return int(load_library().ddpca_admm_launch_count(self._h, C.c_int(1 if reset else 0)))

    def profile(self, enable: bool):
        check(load_library().ddpca_admm_profile(self._h, C.c_int(1 if enable else 0)))

    def profile_get(self, nlevels=16):
        """{(kernel_class, level): (ms, launches, algorithmic_bytes)} of the batched body solves since profile(True)."""
        from .mgpis import KERNEL_CLASSES

        out = {}
        for k, name in enumerate(KERNEL_CLASSES):
            for l in range(nlevels):
                ms, nl, by = C.c_double(), C.c_long(), C.c_double()
                check(load_library().ddpca_admm_profile_get(self._h, C.c_int(k), C.c_int(l), C.byref(ms), C.byref(nl), C.byref(by)))
                if nl.value:
                    out[(name, l)] = (ms.value, nl.value, by.value)
        return out

    def body_iters(self):
        """(number of batched hierarchies, CG iteration count of every body in the last step)."""
        nbt = C.c_int()
        it = (C.c_long * self.nb)()
        check(load_library().ddpca_admm_body_iters(self._h, C.byref(nbt), it))
        return nbt.value, list(it)

    def reset(self, consForc=None):
        """Zero initial state again (MCONTACT.h:875-894); consForc: {v: host vector} new load vectors."""
        lib = load_library()
        for v, f in (consForc or {}).items():
            check(lib.ddpca_admm_set_consforc(self._h, C.c_int(v), C.c_void_p(f) if isinstance(f, int) else _pd(np.ascontiguousarray(f, dtype=np.float64))))
        check(lib.ddpca_admm_reset(self._h))
        self.MULT_MAXI = 1000
        self.moniReco = [[0.0] * 10 for _ in range(self.nb + 4 * self.ni)]
        self.cg_iters = 0
        self.cg_dof_iters = 0.0
