"""Data-parallel gradient exchange: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference has no multi-GPU path; the batch shards naturally, weights are replicated and the one
exchange step is a sum all-reduce of the flat fp32 gradient buffer, issued in buckets so that NCCL pipelines them;
the 1/world scale is folded into the optimiser kernel (mvae_optim_norms grad_scale).  BatchNorm uses per-replica
statistics (what Keras does under MirroredStrategy)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def _prod(shape):
    n = 1
    for v in shape:
        n *= int(v)
    return n


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

    # ---- per-level exchange, launched from inside the step (and captured into its CUDA graph) ---------------------------
    def level_ranges(self, levels):
        """[lo, hi) of the gradient buffer per pyramid level: the variables of encoder_i and of decoder_i are contiguous
        (creation order, engine.Spec.declare_params).  Returns {level: [(lo, hi), (lo, hi)]}; together they tile the
        buffer exactly once."""
        out = {}
        for i in range(levels):
            rs = []
            for side in ("encoder", "decoder"):
                offs = [(e["offset"], e["offset"] + -(-int(_prod(e["shape"])) // 64) * 64) for name, e in self.ps.entries.items()
                        if name.startswith(f"{side}_{i}_")]
                rs.append((min(o[0] for o in offs), max(o[1] for o in offs)))
            out[i] = rs
        flat = sorted(r for rs in out.values() for r in rs)
        assert flat[0][0] == 0 and flat[-1][1] == self.ps.grads.numel() and \
            all(flat[k][1] == flat[k + 1][0] for k in range(len(flat) - 1)), "level ranges do not tile the gradient buffer"
        return out

    def allreduce_level(self, ranges):
        """Sum all-reduce of one level's gradient ranges, asynchronous on the process group's stream: called from the
        level's own stream the moment its backward pass (weight gradients included) is complete, so the exchange of the
        small levels hides under level 0's backward and only level 0's own buckets are exposed.  Returns the work
        handles; `wait_all` joins them into the current stream."""
        if self.world == 1:
            return []
        g = self.ps.grads
        return [dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM, async_op=True) for lo, hi in ranges]

    @staticmethod
    def wait_all(works):
        for w in works:
            w.wait()

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
