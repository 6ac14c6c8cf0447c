"""CPU restatement (numpy / scipy.sparse) of the ADMM loop MCONTACT::CONTACT_ANALYSIS and its
stopping test MCONTACT::MONITOR.  TEST INFRASTRUCTURE ONLY -- the parity checker of the CUDA
path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Parity status: PINNED by tests/test_oracle_golden.py against resuMoni.txt rows, iteration
counts and final states written by the untouched reference (oracle/ref_drivers/block_admm.cpp).

Line references are to /root/reference/MCONTACT.h unless stated otherwise.  All operators are
the reference's own matrices (dumped after MCONTACT::ESTABLISH), so sparsity patterns are
identical by construction; linear solves are exact sparse factorizations (scipy SuperLU)
where the reference uses SimplicialLDLT or MG-PCG to 1e-14.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _csr(d, name):
    shp = d[name + ".shape"]
    return sp.csr_matrix((d[name + ".val"], d[name + ".colidx"], d[name + ".rowptr"]), shape=(int(shp[0]), int(shp[1])))


def vect_medi_osci(v):
    """VECT_MEDI_OSCI, PREP.h:147-153."""
    mx, mn = max(v), min(v)
    return (mx + mn) / 2.0, mx - mn


def contact_projection(g, mu):
    """The contact projection of the interface block, MCONTACT.h:2637-2668, on gamma = 0.5 (trace - gapTerm):
    mu < 0 tied (nothing), mu == 0 frictionless (one component per point, clamp), mu > 0 Coulomb (three components
    per point: normal clamp, tangential pair scaled back onto the cone; status 0 open / 1 slide / 2 stick in the
    second component's slot, as OUTPUT_PRTR writes it, :118).  Returns (gamma, fricStat).
    Pinned to the reference's DEHW run by tests/test_oracle_golden.py (tests/golden/dehw_friction.ddpk.gz)."""
    stat = np.zeros(g.shape[0], dtype=np.int32)
    if mu == 0.0:
        g = np.maximum(0.0, g)  # :2640
    elif mu > 0.0:
        g = g.copy()
        g[0::3] = np.maximum(0.0, g[0::3])  # :2643
        gn = g[0::3]
        t1, t2 = g[1::3].copy(), g[2::3].copy()
        nrm = np.sqrt(t1 * t1 + t2 * t2)
        slid = mu * gn
        open_ = ~(gn > 0.0)
        slide = (gn > 0.0) & (nrm >= slid)  # :2653
        with np.errstate(divide="ignore", invalid="ignore"):
            fac = np.where(slide, slid / nrm, 1.0)
        t1 = np.where(open_, 0.0, t1 * fac)  # :2654-2655, :2663-2664
        t2 = np.where(open_, 0.0, t2 * fac)
        g[1::3], g[2::3] = t1, t2
        stat[1::3] = np.where(open_, 0, np.where(slide, 1, 2))  # :2656-2665
    return g, stat


class AdmmOracle:
    def __init__(self, d: dict):
        self.nb = int(d["nbody"][0])
        self.ni = int(d["niface"][0])
        self.muscSett = int(d["muscSett"][0])
        self.MULT_MAXI = 1000  # PREP.h:75 (mutable global)
        self.body = []
        for v in range(self.nb):
            p = f"body{v}."
            L = int(d[p + "maxiLeve"][0])
            K = _csr(d, p + f"consStif{L}")
            b = {
                "K": K,
                "solve": spla.splu(K.tocsc()).solve,  # mugrDiso / CG_SOLV, :2527-2532
                "consForc": d[p + "consForc"],
                "forcOper": _csr(d, p + "forcOper"),  # ADDITIONAL_FORCE, MULTIGRID.h:1257-1261
                "dispCons": d[p + "dispCons"],  # OUTP_SUB1 constant part, MULTIGRID.h:1272-1279
                "nfull": int(d[p + "nfull"][0]),
            }
            if self.muscSett & 3:
                b["accuProl"] = _csr(d, p + "accuProl")
            if self.muscSett & 2:
                b["globTran_D_1"] = _csr(d, p + "globTran_D_1")  # MCONTACT.h:1868-2055
            self.body.append(b)
        self.iface = []
        for ts in range(self.ni):
            p = f"if{ts}."
            it = {
                "contBody": [int(x) for x in d[p + "contBody"]],
                "fricCoef": float(d[p + "fricCoef"][0]),
                "nip": int(d[p + "nip"][0]),
                "gapTerm": d[p + "gapTerm"],  # pemaInpo * inpoNgap, :2636
                "side": [],
            }
            for tv in range(2):
                q = p + f"s{tv}."
                s = {k: _csr(d, q + k) for k in ("systTran", "systTran_pena", "inteMass", "inteMass_pena", "inpoLagr", "inteInpo", "pemaInpo_r")}
                s["solve_mass"] = spla.splu(s["inteMass"].tocsc()).solve  # inteDiso, :2696
                s["solve_mass_pena"] = spla.splu(s["inteMass_pena"].tocsc()).solve  # inteDiso_pena, :2677
                if self.muscSett & 1:
                    for k in ("globTran", "globTran_pena", "globTran_D"):
                        s[k] = _csr(d, q + k)
                if self.muscSett & 2:
                    s["globTran_1"] = _csr(d, q + "globTran_1")  # :2124-2298
                it["side"].append(s)
            self.iface.append(it)
        if self.muscSett & 3:
            self.baseReco = [int(x) for x in d["baseReco"]]
        if self.muscSett & 1:
            self.globCoup = _csr(d, "globCoup")
            self.glob_solve = spla.splu(self.globCoup.tocsc()).solve  # coarSolv_D, :2553
        if self.muscSett & 2:  # interface-eliminated coarse problem, MULTISCALE_1 (:1672-2343)
            self.globCoup_1 = _csr(d, "globCoup_1")
            self.glob_solve_1 = spla.splu(self.globCoup_1.tocsc()).solve  # coarSolv_D_1, :2588
            self.globForc_1 = d["globForc_1"]
        # state, zero-initialised (:875-894)
        self.resuDisp = [np.zeros(b["nfull"]) for b in self.body]
        self.inteAuxi = [[np.zeros(s["inteMass"].shape[0]) for s in it["side"]] for it in self.iface]
        self.inteLagr = [[np.zeros(s["inteMass"].shape[0]) for s in it["side"]] for it in self.iface]
        self.inpoGamm = [None] * self.ni
        self.fricStat = [None] * self.ni
        self.moniReco = [[0.0] * 10 for _ in range(self.nb + 4 * self.ni)]  # :2494-2498
        self.iterNumbReco = None
        self.rows = []

    # ---------------------------------------------------------------------------------------
    def outp_sub1(self, v, u):
        b = self.body[v]
        return b["forcOper"].T @ u + b["dispCons"]

    def step(self, tc):
        """One pass of the loop body, :2505-2704 (file output omitted)."""
        macro = (self.muscSett & 1) and tc <= self.MULT_MAXI
        # ---- body balance :2511-2538
        for v, b in enumerate(self.body):
            addiForc = np.zeros(b["nfull"])
            for ts, it in enumerate(self.iface):
                for ti in range(2):
                    if it["contBody"][ti] != v:
                        continue
                    s = it["side"][ti]
                    addiForc += s["systTran_pena"] @ self.inteAuxi[ts][ti] - s["systTran"] @ self.inteLagr[ts][ti]  # :2520-2521
            red = b["forcOper"] @ addiForc  # :2524
            u = b["solve"](b["consForc"] + red)  # :2528 / :2531
            self.resuDisp[v] = self.outp_sub1(v, u)  # :2533
        # ---- macroscopic problem :2540-2573
        if macro:
            globForc = np.zeros(self.globCoup.shape[0])
            for ts, it in enumerate(self.iface):
                for tv in range(2):
                    s = it["side"][tv]
                    globForc += s["globTran"] @ self.inteLagr[ts][tv] - s["globTran_pena"] @ self.inteAuxi[ts][tv] + s["globTran_D"] @ self.resuDisp[it["contBody"][tv]]
            globSolu = self.glob_solve(globForc)  # :2553
            for v, b in enumerate(self.body):
                nrow = b["accuProl"].shape[1]
                seg = globSolu[self.baseReco[v] : self.baseReco[v] + nrow]  # :2564-2566
                u = b["accuProl"] @ seg  # :2567
                self.resuDisp[v] = self.resuDisp[v] + (b["forcOper"].T @ u + b["dispCons"])  # :2569-2570 (OUTP_SUB1 re-adds prescribed values)
        # ---- interface-eliminated coarse problem :2575-2607
        if (self.muscSett & 2) and tc <= self.MULT_MAXI:
            globForc = self.globForc_1.copy()  # :2576
            for ts, it in enumerate(self.iface):
                for tv in range(2):
                    globForc += it["side"][tv]["globTran_1"] @ self.inteLagr[ts][tv]  # :2579
            for v, b in enumerate(self.body):
                globForc -= b["globTran_D_1"] @ self.resuDisp[v]  # :2583
            globSolu = self.glob_solve_1(globForc)  # :2588
            for v, b in enumerate(self.body):
                nrow = b["accuProl"].shape[1]
                seg = globSolu[self.baseReco[v] : self.baseReco[v] + nrow]  # :2599-2601
                u = b["accuProl"] @ seg  # :2602
                self.resuDisp[v] = self.resuDisp[v] + (b["forcOper"].T @ u + b["dispCons"])  # :2603-2604
        # ---- interface balance :2628-2685
        for ts, it in enumerate(self.iface):
            s0, s1 = it["side"]
            u0, u1 = self.resuDisp[it["contBody"][0]], self.resuDisp[it["contBody"][1]]
            g = 0.5 * (s0["inpoLagr"] @ self.inteLagr[ts][0] - s1["inpoLagr"] @ self.inteLagr[ts][1] + s0["pemaInpo_r"] @ u0 - s1["pemaInpo_r"] @ u1 - it["gapTerm"])  # :2632-2636
            mu = it["fricCoef"]
            nip = it["nip"]
            g, stat = contact_projection(g, mu)  # :2637-2668
            assert g.shape[0] == (nip if mu == 0.0 else 3 * nip)
            self.inpoGamm[ts], self.fricStat[ts] = g, stat
            for tv, s in enumerate(it["side"]):
                u = self.resuDisp[it["contBody"][tv]]
                inteForc = s["systTran_pena"].T @ u + s["inteMass"] @ self.inteLagr[ts][tv] + s["inteInpo"] @ g  # :2672-2675
                self.inteAuxi[ts][tv] = s["solve_mass_pena"](inteForc)  # :2677
        # ---- Lagrange multiplier :2689-2704
        for ts, it in enumerate(self.iface):
            for tv, s in enumerate(it["side"]):
                u = self.resuDisp[it["contBody"][tv]]
                inteForc = s["systTran_pena"].T @ u - s["inteMass_pena"] @ self.inteAuxi[ts][tv]  # :2692-2694
                self.inteLagr[ts][tv] = self.inteLagr[ts][tv] + s["solve_mass"](inteForc)  # :2696

    def monitor(self, tc, disp0, auxi0, lagr0):
        """MCONTACT::MONITOR, :2725-2845.  Returns (row, flag0, flag1)."""
        cyc = 10
        flag0 = tc >= cyc
        flag1 = True
        convValu = convCrit = 0.0
        row = []
        for v in range(self.nb):
            dv = float(np.sum((self.resuDisp[v] - disp0[v]) ** 2))
            al = float(np.sum(self.resuDisp[v] ** 2))
            self.moniReco[v][tc % cyc] = dv
            convValu += dv
            convCrit += al
            row += [dv, al]
            if tc >= cyc:
                medi, osci = vect_medi_osci(self.moniReco[v])
                if osci > 0.1 * medi:
                    flag0 = False
            if dv > 1.0e-12 * al:
                flag1 = False
        for ts in range(self.ni):
            for tv in range(2):
                k = self.nb + 4 * ts + 2 * tv
                da = float(np.sum((self.inteAuxi[ts][tv] - auxi0[ts][tv]) ** 2))
                aa = float(np.sum(self.inteAuxi[ts][tv] ** 2))
                self.moniReco[k][tc % cyc] = da
                convValu += da
                convCrit += aa
                row += [da, aa]
                if tc >= cyc:
                    medi, osci = vect_medi_osci(self.moniReco[k])
                    if osci > 0.1 * medi:
                        flag0 = False
                if da > 1.0e-12 * aa:
                    flag1 = False
                dl = float(np.sum((self.inteLagr[ts][tv] - lagr0[ts][tv]) ** 2))
                la = float(np.sum(self.inteLagr[ts][tv] ** 2))
                self.moniReco[k + 1][tc % cyc] = dl
                row += [dl, la]  # lambda criteria are computed but disabled, :2822,:2830
        row += [convValu, convCrit]
        return row, flag0, flag1

    def run(self, max_iter=3000):
        """The loop of CONTACT_ANALYSIS, :2504-2712.  Returns iterNumbReco."""
        tc = 0
        while tc < max_iter:
            disp0 = [x.copy() for x in self.resuDisp]
            auxi0 = [[x.copy() for x in s] for s in self.inteAuxi]
            lagr0 = [[x.copy() for x in s] for s in self.inteLagr]
            self.step(tc)
            row, flag0, flag1 = self.monitor(tc, disp0, auxi0, lagr0)
            self.rows.append(row)
            if flag0:
                self.MULT_MAXI = tc  # :2838-2840
            if flag1:
                break  # :2841-2843, :2709-2711
            tc += 1
        self.iterNumbReco = tc
        return tc


class PartitionedAdmmOracle(AdmmOracle):
    """The same loop split over ranks the way the multi-GPU product does it:
    every rank owns some bodies and the interface sides attached to them; per iteration three
    sum-all-reduces couple the ranks -- the coarse right-hand side, the side traces of cross-rank
    interfaces, the MONITOR sums.  `allreduce(ndarray) -> ndarray` is supplied by the caller
    (torch.distributed / gloo in tests).  TEST INFRASTRUCTURE: proves on CPU that the exchanged
    data suffice and that the partitioned iteration is the reference iteration."""

    def __init__(self, d, body_rank, rank, allreduce):
        super().__init__(d)
        self.body_rank = list(body_rank)
        self.rank = rank
        self.allreduce = allreduce
        self.local_body = [r == rank for r in self.body_rank]

    def _side_local(self, ts, tv):
        return self.local_body[self.iface[ts]["contBody"][tv]]

    def step(self, tc):
        macro = (self.muscSett & 1) and tc <= self.MULT_MAXI
        for v, b in enumerate(self.body):
            if not self.local_body[v]:
                continue
            addiForc = np.zeros(b["nfull"])
            for ts, it in enumerate(self.iface):
                for ti in range(2):
                    if it["contBody"][ti] == v:
                        s = it["side"][ti]
                        addiForc += s["systTran_pena"] @ self.inteAuxi[ts][ti] - s["systTran"] @ self.inteLagr[ts][ti]
            u = b["solve"](b["consForc"] + b["forcOper"] @ addiForc)
            self.resuDisp[v] = self.outp_sub1(v, u)
        if macro:
            globForc = np.zeros(self.globCoup.shape[0])
            for ts, it in enumerate(self.iface):
                for tv in range(2):
                    if not self._side_local(ts, tv):
                        continue
                    s = it["side"][tv]
                    globForc += s["globTran"] @ self.inteLagr[ts][tv] - s["globTran_pena"] @ self.inteAuxi[ts][tv] + s["globTran_D"] @ self.resuDisp[it["contBody"][tv]]
            globForc = self.allreduce(globForc)            # exchange 1
            globSolu = self.glob_solve(globForc)            # replicated
            for v, b in enumerate(self.body):
                if not self.local_body[v]:
                    continue
                seg = globSolu[self.baseReco[v] : self.baseReco[v] + b["accuProl"].shape[1]]
                self.resuDisp[v] = self.resuDisp[v] + (b["forcOper"].T @ (b["accuProl"] @ seg) + b["dispCons"])
        if (self.muscSett & 2) and tc <= self.MULT_MAXI:   # interface-eliminated coarse problem, :2575-2607
            part = np.zeros(self.globCoup_1.shape[0])
            for ts, it in enumerate(self.iface):
                for tv in range(2):
                    if self._side_local(ts, tv):
                        part += it["side"][tv]["globTran_1"] @ self.inteLagr[ts][tv]
            for v, b in enumerate(self.body):
                if self.local_body[v]:
                    part -= b["globTran_D_1"] @ self.resuDisp[v]
            globForc = self.allreduce(part) + self.globForc_1   # same exchange buffer; constant part added once
            globSolu = self.glob_solve_1(globForc)               # replicated
            for v, b in enumerate(self.body):
                if not self.local_body[v]:
                    continue
                seg = globSolu[self.baseReco[v] : self.baseReco[v] + b["accuProl"].shape[1]]
                self.resuDisp[v] = self.resuDisp[v] + (b["forcOper"].T @ (b["accuProl"] @ seg) + b["dispCons"])
        # side traces; cross-rank ones go through one packed all-reduce
        traces = {}
        packed, layout = [], []
        for ts, it in enumerate(self.iface):
            cross = self.body_rank[it["contBody"][0]] != self.body_rank[it["contBody"][1]]
            for tv in range(2):
                s = it["side"][tv]
                if self._side_local(ts, tv):
                    t = s["inpoLagr"] @ self.inteLagr[ts][tv] + s["pemaInpo_r"] @ self.resuDisp[it["contBody"][tv]]
                else:
                    t = np.zeros(it["gapTerm"].shape[0])
                traces[(ts, tv)] = t
                if cross:
                    layout.append((ts, tv, len(t)))
                    packed.append(t)
        if packed:
            flat = self.allreduce(np.concatenate(packed))   # exchange 2
            off = 0
            for ts, tv, n in layout:
                traces[(ts, tv)] = flat[off : off + n]
                off += n
        for ts, it in enumerate(self.iface):
            if not (self._side_local(ts, 0) or self._side_local(ts, 1)):
                continue
            g = 0.5 * (traces[(ts, 0)] - traces[(ts, 1)] - it["gapTerm"])
            mu = it["fricCoef"]
            stat = np.zeros(g.shape[0], dtype=np.int32)
            if mu == 0.0:
                g = np.maximum(0.0, g)
            elif mu > 0.0:
                g = g.copy()
                g[0::3] = np.maximum(0.0, g[0::3])
                gn, t1, t2 = g[0::3], g[1::3].copy(), g[2::3].copy()
                nrm = np.sqrt(t1 * t1 + t2 * t2)
                slid = mu * gn
                open_ = ~(gn > 0.0)
                slide = (gn > 0.0) & (nrm >= slid)
                with np.errstate(divide="ignore", invalid="ignore"):
                    fac = np.where(slide, slid / nrm, 1.0)
                g[1::3], g[2::3] = np.where(open_, 0.0, t1 * fac), np.where(open_, 0.0, t2 * fac)
                stat[1::3] = np.where(open_, 0, np.where(slide, 1, 2))
            self.inpoGamm[ts], self.fricStat[ts] = g, stat
            for tv, s in enumerate(it["side"]):
                if not self._side_local(ts, tv):
                    continue
                u = self.resuDisp[it["contBody"][tv]]
                inteForc = s["systTran_pena"].T @ u + s["inteMass"] @ self.inteLagr[ts][tv] + s["inteInpo"] @ g
                self.inteAuxi[ts][tv] = s["solve_mass_pena"](inteForc)
        for ts, it in enumerate(self.iface):
            for tv, s in enumerate(it["side"]):
                if not self._side_local(ts, tv):
                    continue
                u = self.resuDisp[it["contBody"][tv]]
                inteForc = s["systTran_pena"].T @ u - s["inteMass_pena"] @ self.inteAuxi[ts][tv]
                self.inteLagr[ts][tv] = self.inteLagr[ts][tv] + s["solve_mass"](inteForc)

    def monitor(self, tc, disp0, auxi0, lagr0):
        """MONITOR on all-reduced sums: every rank reaches the same decision."""
        sums = np.zeros(2 * (self.nb + 4 * self.ni))
        for v in range(self.nb):
            if self.local_body[v]:
                sums[2 * v] = float(np.sum((self.resuDisp[v] - disp0[v]) ** 2))
                sums[2 * v + 1] = float(np.sum(self.resuDisp[v] ** 2))
        for ts in range(self.ni):
            for tv in range(2):
                if not self._side_local(ts, tv):
                    continue
                k = self.nb + 4 * ts + 2 * tv
                sums[2 * k] = float(np.sum((self.inteAuxi[ts][tv] - auxi0[ts][tv]) ** 2))
                sums[2 * k + 1] = float(np.sum(self.inteAuxi[ts][tv] ** 2))
                sums[2 * (k + 1)] = float(np.sum((self.inteLagr[ts][tv] - lagr0[ts][tv]) ** 2))
                sums[2 * (k + 1) + 1] = float(np.sum(self.inteLagr[ts][tv] ** 2))
        sums = self.allreduce(sums)                          # exchange 3
        cyc = 10
        flag0, flag1 = tc >= cyc, True
        convValu = convCrit = 0.0
        row = []
        for v in range(self.nb):
            dv, al = sums[2 * v], sums[2 * v + 1]
            self.moniReco[v][tc % cyc] = dv
            convValu += dv
            convCrit += al
            row += [dv, al]
            if tc >= cyc:
                medi, osci = vect_medi_osci(self.moniReco[v])
                if osci > 0.1 * medi:
                    flag0 = False
            if dv > 1.0e-12 * al:
                flag1 = False
        for ts in range(self.ni):
            for tv in range(2):
                k = self.nb + 4 * ts + 2 * tv
                da, aa, dl, la = sums[2 * k], sums[2 * k + 1], sums[2 * (k + 1)], sums[2 * (k + 1) + 1]
                self.moniReco[k][tc % cyc] = da
                convValu += da
                convCrit += aa
                row += [da, aa, dl, la]
                if tc >= cyc:
                    medi, osci = vect_medi_osci(self.moniReco[k])
                    if osci > 0.1 * medi:
                        flag0 = False
                if da > 1.0e-12 * aa:
                    flag1 = False
                self.moniReco[k + 1][tc % cyc] = dl
        row += [convValu, convCrit]
        return row, flag0, flag1
