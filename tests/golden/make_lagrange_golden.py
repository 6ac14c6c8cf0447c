"""Golden results of the reference's dual-mortar path MCONTACT::LAGRANGE(1) (MCONTACT.h:2847-3701) on CYLINDER_1
(examples/CYLINDER.cpp:85-90, reduced to locaLeve 5): run by the UNTOUCHED reference (oracle/_ref/cylinder_lagrange,
built by oracle/Makefile from /root/reference) in the build container, 3 min.  The JSON holds what the overlay run on
the GPU box is compared with (tools/lagrange_bench.py, tests/test_gpu_overlay.py): active-set steps, BiCGSTAB iteration
counts, displacement norms per body and the multipliers (contact tractions) the reference writes to resuLagr_*.txt
(:3618-3635).
  python tests/golden/make_lagrange_golden.py"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def lagr_summary(cwd, sub):
    """Per resuLagr_<ts>.txt (columns: node, status, normal, tangential 1, tangential 2): rows, how many constraints are
    active in each status, Euclidean norm of the normal multipliers."""
    out = []
    d = os.path.join(cwd, sub)
    names = sorted((n for n in os.listdir(d) if n.startswith("resuLagr_")), key=lambda n: int(n[9:-4]))
    for n in names:
        a = np.loadtxt(os.path.join(d, n), ndmin=2)
        if a.size == 0:
            out.append({"file": n, "rows": 0, "status_counts": {}, "normal_norm": 0.0, "normal_max": 0.0})
            continue
        st, cnt = np.unique(a[:, 1].astype(int), return_counts=True)
        out.append({"file": n, "rows": int(a.shape[0]), "status_counts": {str(int(s)): int(c) for s, c in zip(st, cnt)},
                    "normal_norm": float(np.linalg.norm(a[:, 2])), "normal_max": float(np.abs(a[:, 2]).max())})
    return out


def run(exe, args, sub):
    tmp = tempfile.mkdtemp(prefix="ddpca_lagr_")
    txt = subprocess.check_output([exe] + args, cwd=tmp).decode()
    res = json.loads(txt.strip().splitlines()[-1])
    res["resuLagr"] = lagr_summary(tmp, sub)
    return res


if __name__ == "__main__":
    res = run(os.path.join(ROOT, "oracle", "_ref", "cylinder_lagrange"), ["--loca", "5"], "Cylinder")
    json.dump(res, open(os.path.join(HERE, "cylinder_lagrange.json"), "w"), indent=1)
    print(json.dumps(res)[:600])
