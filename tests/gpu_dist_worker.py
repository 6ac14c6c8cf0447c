"""Worker of the multi-GPU ADMM test (one process per GPU, NCCL).  Usage (under torchrun):
   gpu_dist_worker.py <ddpk file> <out dir> <muscSett> <max iters>"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))

import ddpca_b200 as dd  # noqa: E402
from ddpca_b200 import ddpk  # noqa: E402
from ddpca_b200.comm import TorchComm  # noqa: E402
from ddpca_b200.partition import partition_bodies  # noqa: E402


def main():
    path, out_dir, musc, iters = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    d = ddpk.load(path)
    nb, ni = int(d["nbody"][0]), int(d["niface"][0])
    contBody = [[int(x) for x in d[f"if{ts}.contBody"]] for ts in range(ni)]
    weights = [len(d[f"body{v}.consStif{int(d[f'body{v}.maxiLeve'][0])}.val"]) for v in range(nb)]
    body_rank = partition_bodies(weights, contBody, world)
    comm = TorchComm(torch.device("cuda", local))
    factorize = None
    if "if0.s0.inteDiso.perm" not in d:
        from tests.helpers import dense_ldlt_factor as factorize
    mc = dd.MCONTACT.from_ddpk(d, device=local, muscSett=musc, factorize=factorize, body_rank=body_rank, rank=rank, comm=comm)
    mc.CONTACT_ANALYSIS(maxiIter=iters)
    disp = mc.resuDisp
    res = {"rank": rank, "world": world, "body_rank": body_rank, "iterNumbReco": mc.iterNumbReco,
           "rows": [list(map(float, r)) for r in mc.resuMoni],
           "disp_norm": {str(v): float(np.linalg.norm(disp[v])) for v in range(nb) if body_rank[v] == rank},
           "cg_iters": mc.cg_iters, "launches": mc.launch_count()}
    json.dump(res, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    mc.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
