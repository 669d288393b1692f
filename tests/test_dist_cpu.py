"""Data-parallel gradient exchange on CPU: world_size 2 over gloo (the NCCL path runs the same code on the GPU box)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _PS:
    def __init__(self, n, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self.flat = torch.randn(n, generator=g)
        self.grads = torch.randn(n, generator=g)
        self.acc = torch.full((n,), 0.1) * (rank + 1)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiscale_variational_autoencoder_b200.dist import GradAllReduce
    ps = _PS(n, rank)
    expect = sum(_PS(n, r).grads for r in range(world))
    ar = GradAllReduce(ps, torch.device("cpu"), bucket_mb=0.001)      # ~262 floats per bucket -> many buckets
    assert len(ar.buckets) > 3 and ar.buckets[0][1] == n and ar.buckets[-1][0] == 0
    covered = sorted(ar.buckets)
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    ar.broadcast_params()
    ar.allreduce()
    ok = torch.allclose(ps.grads, expect, atol=1e-6) and torch.equal(ps.flat, _PS(n, 0).flat) \
        and torch.equal(ps.acc, _PS(n, 0).acc)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_levels(rank, world, port, q):
    """Per-level exchange (what the step launches from each level's stream): the ranges tile the gradient buffer, and
    exchanging them level by level equals one all-reduce of the whole buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiscale_variational_autoencoder_b200 import engine
    from multiscale_variational_autoencoder_b200.dist import GradAllReduce
    enc = {"filters": [32, 32], "kernel_size": [(3, 3)] * 2, "strides": [(2, 2), (1, 1)]}
    dec = {k: enc[k][::-1] for k in enc}
    spec = engine.Spec((32, 32, 3), [16, 8, 4], enc, dec, 0.0, 255.0, 0.5, None, 1.0, "no_upsample")
    ps = engine.ParamStore(torch.device("cpu"), seed=7)
    spec.declare_params(ps)
    ps.finalize()
    g = torch.Generator().manual_seed(5 + rank)
    ps.grads = torch.randn(ps.size, generator=g)
    expect = sum(torch.randn(ps.size, generator=torch.Generator().manual_seed(5 + r)) for r in range(world))
    ar = GradAllReduce(ps, torch.device("cpu"))
    ranges = ar.level_ranges(3)
    ok = sorted(ranges) == [0, 1, 2] and all(len(v) == 2 for v in ranges.values())
    e0 = ps.entries["encoder_1_conv_base/kernel"]["offset"]
    ok = ok and ranges[1][0][0] == e0 and ranges[0][0][0] == 0
    # the Dense kernels worth an exchange of their own, and what is left for the end of the step: together the buffer, once
    early = ar.early_ranges(min_bytes=1 << 14)
    ok = ok and "encoder_0_mu_log_var/kernel" in early and all(k.endswith("/kernel") for k in early)
    left = ar.leftover_ranges(list(early.values()))
    cover = sorted(list(early.values()) + left)
    ok = ok and cover[0][0] == 0 and cover[-1][1] == ps.size and all(cover[k][1] == cover[k + 1][0] for k in range(len(cover) - 1))
    ok = ok and ar.leftover_ranges([]) == [(0, ps.size)]
    ok = ok and ar.peer is None                     # CPU tensors: the peer-memory transport is not attempted
    works = []
    for i in (2, 1, 0):                       # the order the levels finish in
        works += ar.allreduce_level(ranges[i])
    ar.wait_all(works)
    ok = ok and torch.allclose(ps.grads, expect, atol=1e-5)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_per_level_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_levels, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1500, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
