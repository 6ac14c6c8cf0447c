"""Diagnostic: device time of each phase of one ADMM iteration (bench.py's BEAM DD workload, or --block G) on one GPU,
plus the CG iteration counts of the bodies of the last batched solve (how much lock-step waiting the batch costs)."""
import argparse
import ctypes as C
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

ap = argparse.ArgumentParser()
ap.add_argument("--glob", type=int, default=3)
ap.add_argument("--doma", default="8,2,1")
ap.add_argument("--divi", default="")
ap.add_argument("--musc", type=int, default=1)
ap.add_argument("--cpu-iters", type=int, default=4)
ap.add_argument("--iters", type=int, default=4)
args = ap.parse_args()
path, meta = bench.generate_workload(args)
d = ddpk.load(path, copy=False)
t0 = time.time()
mc = dd.MCONTACT.from_ddpk(d)
print("upload s", round(time.time() - t0, 2), mc.upload_times, "bodies DOF", sum(mc.body_dof), flush=True)
lib = load_library()
names = {0: "bodies", 1: "macro_partial", 2: "macro_apply", 6: "macro1_partial", 7: "macro1_apply", 3: "traces", 4: "interface", 5: "monitor"}
order = [0] + ([1, 2] if mc.muscSett & 1 else []) + ([6, 7] if mc.muscSett & 2 else []) + [3, 4, 5]
row = np.empty(mc.row_len)
for tc in range(args.iters):
    out = {}
    for ph in order:
        torch.cuda.synchronize()
        t0 = time.time()
        check(lib.ddpca_admm_phase(mc._h, C.c_int(ph)))
        torch.cuda.synchronize()
        out[names[ph]] = round(1e3 * (time.time() - t0), 3)
    it, dofit = C.c_long(), C.c_double()
    check(lib.ddpca_admm_monitor_row(mc._h, row.ctypes.data_as(C.POINTER(C.c_double)), C.byref(it), C.byref(dofit)))
    nbt, its = mc.body_iters()
    print("iteration", tc, "phase ms", out, "| batches", nbt, "CG iterations per body", its, "lock-step efficiency %.3f" % (sum(its) / (len(its) * max(its)) if max(its) else 1.0), flush=True)
