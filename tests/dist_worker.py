"""Worker of the world_size-2 gloo tests (tests/test_dist_cpu.py).  Each rank runs the partitioned
ADMM oracle on its own bodies; the three per-iteration exchanges go through torch.distributed."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))

from ddpca_b200 import ddpk  # noqa: E402
from ddpca_b200.partition import cross_interfaces, partition_bodies  # noqa: E402
from oracle.admm_oracle import PartitionedAdmmOracle  # noqa: E402


def main():
    out_dir, musc, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dump = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "tests", "golden", "block_small.ddpk.gz")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    d = ddpk.load(dump)
    nb, ni = int(d["nbody"][0]), int(d["niface"][0])
    contBody = [[int(x) for x in d[f"if{ts}.contBody"]] for ts in range(ni)]
    weights = [len(d[f"body{v}.consStif{int(d[f'body{v}.maxiLeve'][0])}.val"]) for v in range(nb)]
    body_rank = partition_bodies(weights, contBody, world)
    exchanged = {"n": 0, "bytes": 0}

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        exchanged["n"] += 1
        exchanged["bytes"] += t.numel() * 8
        return t.numpy()

    o = PartitionedAdmmOracle(d, body_rank, rank, allreduce)
    o.muscSett = musc
    it = o.run(max_iter=iters)
    res = {
        "rank": rank, "world": world, "body_rank": body_rank, "iterNumbReco": it, "rows": o.rows,
        "cross": cross_interfaces(contBody, body_rank), "allreduces": exchanged["n"], "bytes": exchanged["bytes"],
        "disp_norm": {str(v): float(np.linalg.norm(o.resuDisp[v])) for v in range(nb) if body_rank[v] == rank},
    }
    json.dump(res, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
