#!/usr/bin/env python
"""Exploration helper (GPU box): run the reference's DEHW set-up once (oracle/_ref/dehw_admm --ref-iters 1),
then iterate the DEVICE loop and print, every few iterations, how many integration points of each frictional
interface are open / sliding / sticking -- used to choose the iteration at which the reference-pinned friction
fixtures and tests look (tests/test_gpu_examples.py, tests/golden/make_dehw_friction_fixture.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ddpca-admm_b200")):
    sys.path.insert(0, p)
import ddpca_b200 as dd  # noqa: E402
from tests.helpers import run_ref_dehw  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
t0 = time.time()
d, meta = run_ref_dehw(1)
print("reference set-up + 1 iteration: %.1f s" % (time.time() - t0), meta["body_dof"], flush=True)
t0 = time.time()
mc = dd.MCONTACT.from_ddpk(d)
print("upload: %.1f s" % (time.time() - t0), flush=True)
t0 = time.time()
for tc in range(iters):
    row = mc.step(tc)
    conv = mc.MONITOR(tc, row)
    if tc % 10 == 9 or conv == 1:
        out = []
        for ts in range(mc.ni):
            if mc.fricCoef[ts] > 0:
                g, st = mc.inpoGamm(ts)
                c = np.bincount(st[1::3], minlength=3)
                out.append("if%d: open %d slide %d stick %d" % (ts, c[0], c[1], c[2]))
        print("it %d (%.1f s) Cvalu %.3e Ccrit %.3e MULT_MAXI %d | %s" % (tc, time.time() - t0, row[-2], row[-1], mc.MULT_MAXI, " | ".join(out)), flush=True)
    if conv == 1:
        print("converged at", tc)
        break
mc.close()
