"""Batch-sharded data parallelism for the AGCN model: one process per GPU, full replica per rank, one bucketed
gradient all-reduce per step (NCCL over NVLink 5 / NVSwitch on the GPU box; gloo in the CPU tests).

The reference has no distributed code at all; the only natural shard of the path is the batch
(SURVEY 8e).  BatchNorm statistics stay per replica, which is what DistributedDataParallel does by default.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_batch(n_global: int, rank: int, world: int):
    """Rank r takes samples [r*N/R, (r+1)*N/R)."""
    if n_global % world:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return rank * per, (rank + 1) * per


class GradientAllReducer:
    """Averages parameter gradients across ranks in fixed-size flat buckets.  Buckets are filled in reverse
    parameter order (the order backward produces them) and reduced asynchronously; ``wait()`` blocks the
    current stream on the communication and scatters the averaged values back into ``.grad``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 8 << 20, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._pending = []

    def start(self):
        if self.world == 1:
            return
        for bucket in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, flat, bucket))

    def wait(self):
        for work, flat, bucket in self._pending:
            work.wait()
            flat.div_(self.world)
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._pending = []

    def __call__(self):
        self.start()
        self.wait()
