#!/usr/bin/env python
"""BLOCK contact example (examples/BLOCK.h: 3 blocks + 6 plates, 2 frictionless contact + 6 tied interfaces, macroscopic
problem) at globLeve G on one GPU, against the untouched reference run to convergence on this box's host cores
(round-1 yardstick: globLeve 3 took 24.5 s upload + 0.69 s solve, 5 500 launches per iteration, 316 M DOF*iter/s in-loop).
   usage: python tools/block_admm_bench.py [G]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ddpca-admm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import ddpca_b200 as dd  # noqa: E402
from tests.helpers import run_ref_driver  # noqa: E402

glob = int(sys.argv[1]) if len(sys.argv) > 1 else 3
t0 = time.time()
d, meta = run_ref_driver("block_admm", ["--glob", glob, "--musc", 1, "--ref-iters", 0])
t_ref = time.time() - t0
t0 = time.time()
mc = dd.MCONTACT.from_ddpk(d)
torch.cuda.synchronize()
upload_s = time.time() - t0
mc.CONTACT_ANALYSIS()          # warm-up
times = []
for _ in range(3):
    mc.reset()
    mc.launch_count(reset=True)
    torch.cuda.synchronize()
    t0 = time.time()
    mc.CONTACT_ANALYSIS()
    torch.cuda.synchronize()
    times.append(time.time() - t0)
its = mc.iterNumbReco + 1
err = max(float(np.linalg.norm(mc.resuDisp[v] - d[f"ref.resuDisp{v}"]) / np.linalg.norm(d[f"ref.resuDisp{v}"])) for v in range(mc.nb))
solve = min(times)
print(json.dumps({
    "workload": f"BLOCK domaNumb=1x1x1 globLeve={glob}", "body_dof": meta["body_dof"], "globCoup_rows": meta.get("globCoup_rows"),
    "admm_iterations": its, "reference_admm_iterations": meta["ref_iterNumbReco"] + 1,
    "solve_wall_s": solve, "upload_s": round(upload_s, 2), "upload_breakdown_s": mc.upload_times,
    "gpu_launches_per_admm_iteration": round(mc.launch_count() / its, 1), "batches": mc.body_iters()[0],
    "mgpcg_dof_iter_per_s": mc.cg_dof_iters / solve, "cg_iterations_per_solve": mc.cg_iters,
    "max_rel_err_resuDisp_vs_reference": err,
    "reference": {"solve_wall_s": meta["ref_admm_s"], "cores": meta.get("omp_max_threads"), "driver_wall_s": round(t_ref, 1)},
}))
mc.close()
