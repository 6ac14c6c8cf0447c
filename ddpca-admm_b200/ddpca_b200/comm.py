"""Exchange plumbing for the multi-GPU ADMM loop: torch.distributed (NCCL over NVLink) moves three
small device buffers per iteration -- an all-reduce of the coarse right-hand side, a PAIRWISE swap of the signed
interface side traces between the two owners of every cross-rank interface, and an all-reduce of the MONITOR sums.
torch is used for device memory, the stream and the collectives only."""
from __future__ import annotations


class TorchComm:
    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.device = device
        self.rank = dist.get_rank(group)
        # the library and the collectives share one non-null stream: phases and exchanges are ordered on it
        if device.type == "cuda":
            self.stream = torch.cuda.Stream(device=device)
            torch.cuda.set_stream(self.stream)
        else:
            self.stream = None

    def stream_ptr(self) -> int:
        return self.stream.cuda_stream if self.stream is not None else 0

    def alloc(self, nglob: int, ntrace: int, nmoni: int):
        t = self.torch
        mk = lambda n: t.zeros(max(int(n), 1), dtype=t.float64, device=self.device)[: int(n)]
        return mk(nglob), mk(ntrace), mk(ntrace), mk(nmoni)

    def allreduce_sum(self, tensor):
        if tensor is None or tensor.numel() == 0:
            return
        self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)

    def swap(self, send, recv, peers):
        """peers = [(rank, offset, count)]: send[offset:offset+count] goes to `rank`, the same range of recv comes
        from it.  One grouped NCCL call (ncclGroupStart/End underneath), no reduction, no zero padding."""
        d = self.dist
        ops = []
        for rank, off, cnt in peers:
            if cnt == 0:
                continue
            ops.append(d.P2POp(d.isend, send[off:off + cnt], rank, self.group))
            ops.append(d.P2POp(d.irecv, recv[off:off + cnt], rank, self.group))
        if not ops:
            return
        for w in d.batch_isend_irecv(ops):
            w.wait()
