#!/usr/bin/env python
"""Builds tests/golden/dehw_friction.ddpk.gz from a dump of the UNTOUCHED reference's DEHW example.

    oracle/_ref/dehw_admm --dd 0 --homo 1 --loca 1 --ref-iters 260 --nomat --out dehw260.ddpk     (~55 min on 2 cores)
    python tests/golden/make_dehw_friction_fixture.py dehw260.ddpk

DEHW is the only example with frictional interfaces (mu = 0.08, examples/DEHW.h:1619) and its tooth flanks only
touch after ~175 ADMM iterations of 12 s each, so the reference cannot be re-run inside a test.  The dump holds the
reference's state after K = 260 passes of the loop body (resuDisp, inteAuxi, inteLagr), the interface operators and
the resuCont_<ts>.txt files of pass K-1 (normal pressure, tangential traction, Coulomb status per integration point,
MCONTACT.h:97-123).  From the state the multipliers BEFORE the last update follow exactly,
    lambda_{K-1} = lambda_K - M^-1 (S_p^T u_K - M_p aux_K)                                   (MCONTACT.h:2691-2697)
and with them the trace and the projection of pass K-1 (MCONTACT.h:2632-2668) can be replayed and compared with what
the reference wrote.  The fixture keeps, for every frictional interface, a sample of its integration points (all that
slide or stick + as many open ones) and only the rows / columns of the operators those points and the side's
interface nodes touch."""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))
from ddpca_b200 import ddpk  # noqa: E402


def main():
    src = sys.argv[1]
    d = ddpk.load(src, copy=False)
    out = {}
    ni = int(d["niface"][0])
    keep_if = [ts for ts in range(ni) if float(d[f"if{ts}.fricCoef"][0]) > 0.0]
    out["niface"] = np.array([len(keep_if)], dtype=np.int64)
    out["ref_iterations"] = np.array([int(d["ref.first_iters"][0])], dtype=np.int64)
    rng = np.random.default_rng(0)
    for k, ts in enumerate(keep_if):
        p, q = f"if{ts}.", f"f{k}."
        nip = int(d[p + "nip"][0])
        cont = d[f"ref.resuCont{ts}"].reshape(nip, 5)
        active = np.nonzero(cont[:, 4] != 0)[0]
        rest = np.setdiff1d(np.arange(nip), active)
        extra = rng.choice(rest, size=min(len(rest), max(64, len(active))), replace=False)
        ips = np.sort(np.concatenate([active, extra])).astype(np.int64)
        rows = (3 * ips[:, None] + np.arange(3)[None, :]).ravel()          # d = 3 components per point
        out[q + "fricCoef"] = np.array([float(d[p + "fricCoef"][0])])
        out[q + "source_interface"] = np.array([ts], dtype=np.int64)
        out[q + "points"] = ips
        out[q + "gapTerm"] = np.ascontiguousarray(d[p + "gapTerm"][rows])
        out[q + "resuCont"] = np.ascontiguousarray(cont[ips]).ravel()
        cb = [int(x) for x in d[p + "contBody"]]
        for tv in range(2):
            s, t = p + f"s{tv}.", q + f"s{tv}."
            u = d[f"ref.resuDisp{cb[tv]}"]
            Pr = ddpk.get_csr(d, s + "pemaInpo_r").to_scipy()[rows]
            Sp = ddpk.get_csr(d, s + "systTran_pena").to_scipy()
            used = np.union1d(np.unique(Pr.indices), np.nonzero(np.diff(Sp.indptr))[0])     # displacement DOFs that matter
            ddpk.put_csr(out, t + "pemaInpo_r", ddpk.Csr.from_scipy(Pr[:, used]))
            ddpk.put_csr(out, t + "systTran_pena", ddpk.Csr.from_scipy(Sp[used]))
            ddpk.put_csr(out, t + "inpoLagr", ddpk.Csr.from_scipy(ddpk.get_csr(d, s + "inpoLagr").to_scipy()[rows]))
            for name in ("inteMass", "inteMass_pena"):
                ddpk.put_csr(out, t + name, ddpk.get_csr(d, s + name))
            out[t + "resuDisp_used"] = np.ascontiguousarray(u[used])
            out[t + "inteLagr"] = np.ascontiguousarray(d[f"ref.{s}inteLagr"])
            out[t + "inteAuxi"] = np.ascontiguousarray(d[f"ref.{s}inteAuxi"])
    path = os.path.join(ROOT, "tests", "golden", "dehw_friction.ddpk")
    ddpk.save(path, out)
    with open(path, "rb") as f, gzip.open(path + ".gz", "wb", compresslevel=9) as g:
        g.write(f.read())
    os.remove(path)
    print(path + ".gz", os.path.getsize(path + ".gz"), "bytes;", {k: int(len(out[f'f{k}.points'])) for k in range(len(keep_if))}, "points per interface")


if __name__ == "__main__":
    main()
