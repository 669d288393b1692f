"""Data-parallel gradient exchange: one process per GPU over NVLink 5 / NVSwitch.

The reference has no multi-GPU path; the batch shards naturally, weights are replicated and the one
exchange step is a sum all-reduce of the flat fp32 gradient buffer; the 1/world scale is folded into the optimiser
kernel (mvae_optim_norms grad_scale).  BatchNorm uses per-replica statistics (what Keras does under MirroredStrategy).

Two transports:
  * "peer" (default on one node, world <= 8): the library's own kernel over peer memory (csrc/comm.cu, mvae_comm_*): every
    rank maps the peers' gradient buffers through CUDA IPC once, and a captured kernel does barrier -> (rank r pulls and sums
    slice r, pushes the sum to every rank) -> barrier.  torch.distributed only carries the IPC handles at set-up.
  * "nccl": torch.distributed all-reduce in buckets (any world size / several nodes; MVAE_DP_COMM=nccl forces it)."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib


# CUDA IPC handles already opened by this process: handle bytes -> [mapped base, users].  An allocation may only be mapped
# once per process, and the torch allocator can place the buffers of two models in one allocation.
_OPEN = {}


def _ipc_open(lib, handle, offset):
    ent = _OPEN.get(handle)
    if ent is None:
        base, ptr = C.c_void_p(), C.c_void_p()
        _lib.check(lib.mvae_comm_open(handle, 0, C.byref(base), C.byref(ptr)), "mvae_comm_open")
        ent = _OPEN[handle] = [base.value, 0]
    ent[1] += 1
    return ent[0] + offset


def _ipc_close(lib, handle):
    ent = _OPEN.get(handle)
    if ent is None:
        return
    ent[1] -= 1
    if ent[1] <= 0:
        lib.mvae_comm_close(ent[0])
        del _OPEN[handle]


class PeerAllReduce:
    """In-place sum all-reduce of one device buffer over peer memory.  Collective constructor (every rank of the default
    process group must call it with a buffer of the same length)."""

    def __init__(self, buf, device):
        lib = _lib.load()
        self.lib, self.buf, self.device = lib, buf, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if not (2 <= self.world <= 8):
            raise _lib.MvaeError(f"peer all-reduce supports 2..8 ranks, not {self.world}")
        if buf.dtype != torch.float32 or not buf.is_contiguous() or buf.numel() % 4 or buf.data_ptr() % 16:
            raise _lib.MvaeError("peer all-reduce needs a contiguous, 16-byte aligned fp32 buffer of a multiple of 4 elements")
        hb = int(lib.mvae_comm_handle_bytes())
        self._mapped, self._sig = [], None
        # Every step below is local, followed by an exchange of (payload | error): the constructor raises on ALL ranks or on
        # none, so the callers' collectives stay matched.
        mine, err = [], None
        try:
            with torch.cuda.device(device):
                sig = C.c_void_p()
                _lib.check(lib.mvae_comm_alloc_signals(C.byref(sig)), "mvae_comm_alloc_signals")
                self._sig = sig.value
                for ptr in (buf.data_ptr(), self._sig):
                    h = C.create_string_buffer(hb)
                    off = C.c_ulonglong()
                    _lib.check(lib.mvae_comm_export(ptr, h, C.byref(off)), "mvae_comm_export")
                    mine.append((bytes(h.raw), int(off.value)))
        except Exception as e:      # noqa: BLE001
            err = repr(e)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (err, mine, int(buf.numel()), os.uname().nodename))
        self._raise_together([e[0] for e in everyone])
        err = None
        if any(e[2] != buf.numel() for e in everyone):
            err = "the ranks hold gradient buffers of different lengths"
        elif any(e[3] != everyone[0][3] for e in everyone):
            err = "the ranks are not on one node"
        bufs, sigs = (C.c_void_p * self.world)(), (C.c_void_p * self.world)()
        if err is None:
            try:
                with torch.cuda.device(device):
                    for r, (_, handles, _, _) in enumerate(everyone):
                        if r == self.rank:
                            bufs[r], sigs[r] = buf.data_ptr(), self._sig
                            continue
                        for k, (h, off) in enumerate(handles):
                            (bufs if k == 0 else sigs)[r] = _ipc_open(lib, h, off)
                            self._mapped.append(h)
            except Exception as e:      # noqa: BLE001
                err = repr(e)
        errs = [None] * self.world
        dist.all_gather_object(errs, err)       # also: every rank has mapped every peer before the first kernel runs
        self._raise_together(errs)
        self._bufs, self._sigs = bufs, sigs

    def _raise_together(self, errs):
        bad = [(r, e) for r, e in enumerate(errs) if e]
        if bad:
            for h in self._mapped:
                _ipc_close(self.lib, h)
            if self._sig is not None:
                self.lib.mvae_comm_free_signals(self._sig)
            self._mapped, self._sig = [], None
            raise _lib.MvaeError(f"peer all-reduce set-up failed on rank {bad[0][0]}: {bad[0][1]}")

    MAX_RANGES, CHANNELS = 8, 16

    def allreduce(self, stream=None, ctas=0, ranges=None, channel=0):
        """Enqueue the exchange of the element ranges [(lo, hi), ...] (default: the whole buffer; multiples of 4) on `stream`
        (default: the current stream); capturable into a CUDA graph.  Exchanges on one channel must be ordered by stream
        dependencies; different channels may overlap in time."""
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        ranges = [(0, self.buf.numel())] if ranges is None else list(ranges)
        for lo, hi in ranges:
            if lo % 4 or hi % 4 or not (0 <= lo < hi <= self.buf.numel()):
                raise _lib.MvaeError(f"peer all-reduce: bad range [{lo}, {hi})")
        for k in range(0, len(ranges), self.MAX_RANGES):
            part = ranges[k:k + self.MAX_RANGES]
            lo = (C.c_longlong * len(part))(*[r[0] for r in part])
            n = (C.c_longlong * len(part))(*[r[1] - r[0] for r in part])
            _lib.check(self.lib.mvae_comm_allreduce(self._bufs, self._sigs, self.rank, self.world, len(part), lo, n, channel,
                                                    ctas, s), "mvae_comm_allreduce")

    def timed_out(self):
        t = C.c_int()
        _lib.check(self.lib.mvae_comm_status(self._sig, C.byref(t)), "mvae_comm_status")
        return bool(t.value)

    def close(self):
        """Collective: unmap the peers (after everyone has stopped using them) and free the signal block."""
        if self._sig is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()
        for h in self._mapped:
            _ipc_close(self.lib, h)
        self.lib.mvae_comm_free_signals(self._sig)
        self._mapped, self._sig = [], None

    def __del__(self):
        # not collective: only this rank's view of the peers is dropped; the signal block stays allocated (peers may still
        # be inside a kernel that touches it)
        try:
            for h in self._mapped:
                _ipc_close(self.lib, h)
            self._mapped = []
        except Exception:       # noqa: BLE001  (interpreter shutdown)
            pass


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
        # transport: the library's peer-memory kernel when the ranks share a node (CUDA tensors => NCCL group => GPUs),
        # NCCL otherwise or on request.  The choice is collective: a rank that cannot set the peer path up makes all fall back.
        self.peer = None
        want = os.environ.get("MVAE_DP_COMM", "peer")
        if want == "peer" and 2 <= self.world <= 8 and ps.grads.is_cuda:
            try:
                self.peer = PeerAllReduce(ps.grads, device)      # raises on every rank or on none
            except _lib.MvaeError as e:
                if self.rank == 0:
                    import logging
                    logging.getLogger("mvae").warning("%s: gradients go over NCCL instead", e)
        per = max(int(bucket_mb * (1 << 20) / 4), 1)
        # buckets in REVERSE storage order: the decoders (created last) finish their backward first
        self.buckets = []
        hi = n
        while hi > 0:
            lo = max(hi - per, 0)
            self.buckets.append((lo, hi))
            hi = lo

    def early_ranges(self, min_bytes=1 << 20):
        """{weight name: (lo, hi)} of the Dense kernels whose gradient is worth exchanging on its own, as soon as its
        weight-gradient launch has been issued: the Dense heads hold most of the parameters (cfg2: 12.6 of 15.4 MB in two
        variables of level 0) and their gradients are complete long before the level's convolution chain is."""
        out = {}
        for name, e in self.ps.entries.items():
            n = _prod(e["shape"])
            if e["trainable"] and len(e["shape"]) == 2 and 4 * n >= min_bytes and n % 4 == 0 and e["offset"] % 4 == 0:
                out[name] = (e["offset"], e["offset"] + n)
        return out

    def leftover_ranges(self, done):
        """The gradient buffer minus the ranges in `done`, as sorted (lo, hi) pairs."""
        out, pos = [], 0
        for lo, hi in sorted(done):
            if lo > pos:
                out.append((pos, lo))
            pos = max(pos, hi)
        if pos < self.ps.grads.numel():
            out.append((pos, self.ps.grads.numel()))
        return out

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
        if self.peer is not None:
            self.peer.allreduce()
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
