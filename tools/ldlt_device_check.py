#!/usr/bin/env python
"""Debug helper (>= 2 GPUs): the device solve phase of a host SimplicialLDLT factor (ddpca_ldlt_*) on every visible
device -- residual of coarSolv_D_1 of the BEAM DD example (sparse path: more than 4096 rows)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ddpca-admm_b200")):
    sys.path.insert(0, p)
import ddpca_b200 as dd  # noqa: E402
from ddpca_b200 import ddpk  # noqa: E402
from tests.helpers import run_ref_beam_dd  # noqa: E402

d, meta = run_ref_beam_dd(1, doma=(8, 1, 1), musc=2)
A = ddpk.get_csr(d, "globCoup_1").to_scipy()
print("globCoup_1 rows", A.shape[0], "factor in dump:", "coarSolv_D_1.perm" in d)
b = np.random.default_rng(0).standard_normal(A.shape[0])
for dev in range(dd.device_count()):
    s = dd.DIRE_SOLV(d["coarSolv_D_1.perm"], ddpk.get_csr(d, "coarSolv_D_1.L"), d["coarSolv_D_1.D"], device=dev)
    x = s.solve(b)
    print("device", dev, s.info(), "relative residual", float(np.linalg.norm(A @ x - b) / np.linalg.norm(b)))
    s.close()
