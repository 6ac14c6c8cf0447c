"""Worker of the gloo test of ddpca_b200.comm.TorchComm (tests/test_dist_cpu.py): the pairwise trace swap and the
all-reduce the multi-GPU ADMM loop uses, on CPU tensors.  Usage (torchrun): dist_comm_worker.py <out dir>"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ddpca-admm_b200"))

from ddpca_b200.comm import TorchComm  # noqa: E402


def main():
    out_dir = sys.argv[1]
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = TorchComm(torch.device("cpu"))
    # every pair of ranks shares a range of (3 + a + b) doubles, ranges ordered by peer rank like ddpca_admm_exchange_peers
    peers, off = [], 0
    for p in range(world):
        if p == rank:
            continue
        cnt = 3 + min(p, rank) + max(p, rank)
        peers.append((p, off, cnt))
        off += cnt
    send = torch.arange(off, dtype=torch.float64) + 1000.0 * rank
    recv = torch.zeros(off, dtype=torch.float64)
    comm.swap(send, recv, peers)
    ok = True
    for p, o, c in peers:
        # what peer p sent to me sits in ITS range for peer `rank`
        po = 0
        for q in range(world):
            if q == p:
                continue
            cq = 3 + min(q, p) + max(q, p)
            if q == rank:
                break
            po += cq
        expect = torch.arange(po, po + c, dtype=torch.float64) + 1000.0 * p
        ok = ok and bool(torch.equal(recv[o:o + c], expect))
    s = torch.full((5,), float(rank + 1), dtype=torch.float64)
    comm.allreduce_sum(s)
    ok = ok and bool(torch.equal(s, torch.full((5,), world * (world + 1) / 2.0, dtype=torch.float64)))
    json.dump({"rank": rank, "ok": ok, "stream_ptr": comm.stream_ptr()}, open(os.path.join(out_dir, f"comm{rank}.json"), "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
