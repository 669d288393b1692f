"""Data-parallel gradient exchange: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference has no multi-GPU path; the batch shards naturally, weights are replicated and the one
exchange step is a sum all-reduce of the flat fp32 gradient buffer, issued in buckets so that NCCL pipelines them;
the 1/world scale is folded into the optimiser kernel (mvae_optim_norms grad_scale).  BatchNorm uses per-replica
statistics (what Keras does under MirroredStrategy)."""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradAllReduce:
    def __init__(self, ps, device, bucket_mb=32.0):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.ps, self.device = ps, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        n = ps.grads.numel()
        per = max(int(bucket_mb * (1 << 20) / 4), 1)
        # buckets in REVERSE storage order: the decoders (created last) finish their backward first
        self.buckets = []
        hi = n
        while hi > 0:
            lo = max(hi - per, 0)
            self.buckets.append((lo, hi))
            hi = lo

    def broadcast_params(self):
        dist.broadcast(self.ps.flat, src=0)
        if self.ps.acc is not None:
            dist.broadcast(self.ps.acc, src=0)

    def allreduce(self):
        if self.world == 1:
            return
        g = self.ps.grads
        works = [dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, async_op=True) for lo, hi in self.buckets]
        for w in works:
            w.wait()

    def average_moving_stats(self):
        """BatchNorm moving statistics are per replica (every rank normalises with its own batch statistics); before the
        weights are saved they are averaged so that the checkpoint does not depend on which rank writes it."""
        if self.world == 1:
            return
        for name, e in self.ps.entries.items():
            if not e["trainable"]:
                v = self.ps.view(name)
                dist.all_reduce(v, op=dist.ReduceOp.SUM)
                v.div_(self.world)
