"""Exchange plumbing for the multi-GPU ADMM loop: torch.distributed (NCCL over NVLink) all-reduces
of three small device buffers per iteration -- coarse right-hand side, interface side traces,
MONITOR sums.  torch is used for device memory, the stream and the collective only."""
from __future__ import annotations


class TorchComm:
    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.device = device
        # the library and NCCL share one non-null stream: the phases and the all-reduces are ordered on it
        self.stream = torch.cuda.Stream(device=device)
        torch.cuda.set_stream(self.stream)

    def stream_ptr(self) -> int:
        return self.stream.cuda_stream

    def alloc(self, nglob: int, ntrace: int, nmoni: int):
        t = self.torch
        mk = lambda n: t.zeros(max(int(n), 1), dtype=t.float64, device=self.device)[: int(n)]
        return mk(nglob), mk(ntrace), mk(nmoni)

    def allreduce_sum(self, tensor):
        if tensor is None or tensor.numel() == 0:
            return
        with self.torch.cuda.stream(self.stream):
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
