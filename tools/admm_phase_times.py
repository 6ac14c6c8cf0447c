"""Diagnostic: wall time of each phase of one ADMM iteration (BLOCK example) on one GPU."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))
import numpy as np
import torch

import bench
import ddpca_b200 as dd
from ddpca_b200 import ddpk
from ddpca_b200.lib import check, load_library

glob = int(sys.argv[1]) if len(sys.argv) > 1 else 1
path, meta = bench.generate_admm_workload(glob, "")
d = ddpk.load(path)
mc = dd.MCONTACT.from_ddpk(d)
lib = load_library()
names = ["bodies", "macro_partial", "macro_apply", "traces", "interface", "monitor"]
for tc in range(3):
    out = {}
    for ph in range(6):
        torch.cuda.synchronize()
        t0 = time.time()
        check(lib.ddpca_admm_phase(mc._h, C.c_int(ph)))
        torch.cuda.synchronize()
        out[names[ph]] = round(1e3 * (time.time() - t0), 3)
    print("iteration", tc, "phase ms", out, flush=True)
for nm in ("coarSolv_D",):
    s = dd.DIRE_SOLV(d[nm + ".perm"], ddpk.get_csr(d, nm + ".L"), d[nm + ".D"])
    b = np.random.default_rng(0).standard_normal(s.n)
    s.solve(b)
    t0 = time.time()
    for _ in range(3):
        x = s.solve(b)
    print(nm, s.info(), "solve ms", round(1e3 * (time.time() - t0) / 3, 3), "resid", float(np.linalg.norm(ddpk.get_csr(d, "globCoup").to_scipy() @ x - b) / np.linalg.norm(b)))
