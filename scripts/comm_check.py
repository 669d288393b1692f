"""Peer-memory gradient exchange (csrc/comm.cu) on N GPUs of one node, under torchrun:
(1) mvae_comm_allreduce against a float64 sum of the ranks' buffers and against NCCL, bit-identical on all ranks, at lengths
    that do and do not divide by the world size, many times in a row and replayed from a CUDA graph;
(2) device time of one exchange, peer kernel against NCCL, at the gradient sizes of cfg2 / cfg3;
(3) the cfg2 training step with either transport (STEP=1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/comm_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from multiscale_variational_autoencoder_b200.dist import PeerAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def fill(n, r, seed):
    return torch.randn(n, generator=torch.Generator().manual_seed(seed * 100 + r))


# ---- (1) correctness ------------------------------------------------------------------------------------------------
for n in (4, 64, 4 * 1237, 3_850_240, 1 << 20):
    buf = torch.empty(n, device=dev)
    pa = PeerAllReduce(buf, dev)
    for it in range(5):
        buf.copy_(fill(n, rank, it))
        expect = sum(fill(n, r, it).double() for r in range(world))
        pa.allreduce()
        torch.cuda.synchronize()
        err = float((buf.double().cpu() - expect).abs().max())
        assert err <= 1e-5 * max(1.0, float(expect.abs().max())), (n, it, err)
        same = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(same, buf)
        assert all(torch.equal(same[0], s) for s in same), "ranks hold different sums"
    if n >= 4 * 1237:
        # several ranges in one launch on another channel: only they are summed
        rs = [(0, 400), (1000, 1000 + 4 * 333), (n - 4 * 77, n)]
        buf.copy_(fill(n, rank, 55))
        pa.allreduce(ranges=rs, channel=3, ctas=16)
        torch.cuda.synchronize()
        expect = fill(n, rank, 55).double()
        for lo, hi in rs:
            expect[lo:hi] = sum(fill(n, r, 55).double()[lo:hi] for r in range(world))
        err = float((buf.double().cpu() - expect).abs().max())
        assert err <= 1e-5 * float(expect.abs().max()), (n, "ranges", err)
    # replayed from a graph, back to back (the barrier counters advance in device memory)
    buf.copy_(fill(n, rank, 77) * 1e-3)
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        pa.allreduce()
        st.synchronize()
        buf.copy_(fill(n, rank, 78) * 1e-3)
        st.synchronize()
        dist.barrier()
        with torch.cuda.graph(g, stream=st):
            pa.allreduce()
            pa.allreduce()
        g.replay()
        g.replay()
        st.synchronize()
    expect = sum(fill(n, r, 78).double() * 1e-3 for r in range(world)) * world ** 3
    err = float((buf.double().cpu() - expect).abs().max()) / float(expect.abs().max())
    assert err <= 1e-5, (n, "graph", err)
    assert not pa.timed_out()
    pa.close()
    say(f"peer all-reduce n={n}: ok on {world} ranks (eager x5, graph replay x4)")

# ---- (2) device time of one exchange --------------------------------------------------------------------------------
def timed(fn, n_inner=20, reps=5):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    best = 1e9
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        st.synchronize()
        dist.barrier()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(n_inner):
                fn()
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) / n_inner * 1e3)
    t = torch.tensor([best], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for label, n in (("cfg2 gradients", 3_850_240), ("cfg3 gradients", 14_800_000 // 64 * 64), ("256 MB", 64 << 20)):
    buf = torch.zeros(n, device=dev)
    pa = PeerAllReduce(buf, dev)
    for ctas in (32, 64, 128, 256):
        us = timed(lambda: pa.allreduce(ctas=ctas))
        say(f"{label} ({n * 4 / 1e6:.1f} MB) x{world}: peer kernel, {ctas:3d} CTAs: {us:8.1f} us  "
            f"({2 * (world - 1) / world * n * 4 / us / 1e3:.0f} GB/s bus)")
    us = timed(lambda: dist.all_reduce(buf))
    say(f"{label} ({n * 4 / 1e6:.1f} MB) x{world}: NCCL all_reduce:          {us:8.1f} us")
    assert not pa.timed_out()
    pa.close()

# ---- (3) the training step -------------------------------------------------------------------------------------------
if os.environ.get("STEP", "1") == "1":
    import bench
    from multiscale_variational_autoencoder_b200 import MultiscaleVAE
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cfg, B, _ = bench.CONFIGS[name]
    g = torch.Generator().manual_seed(1000 + rank)
    x = torch.rand(B, *cfg["input_dims"], generator=g) * 255
    eps = [torch.randn(B, z, generator=g) for z in cfg["z_dims"]]
    finals = {}
    for comm in ("nccl", "peer-late", "peer"):
        os.environ["MVAE_DP_COMM"] = comm.split("-")[0]
        os.environ["MVAE_DP_EARLY"] = "0" if comm.endswith("late") else "1"
        m = MultiscaleVAE(**cfg, precision="tf32", device=dev, seed=7)
        m.compile(0.01, 1.0, 0.1)
        m.enable_data_parallel()
        assert (m._dist.peer is not None) == comm.startswith("peer")
        eng = m._engine(B, True)
        m._load_input(eng, x.numpy())
        m._load_eps(eng, eps)
        for _ in range(10):
            m.train_step_device(eng)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            m.train_step_device(eng)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # replicas must stay identical: every rank applied the same summed gradient with the same clip factors (the
        # BatchNorm moving statistics are per replica by design and are left out)
        keep = torch.zeros_like(m._ps.flat, dtype=torch.bool)
        for e in m._ps.entries.values():
            if e["trainable"]:
                n_e = 1
                for d in e["shape"]:
                    n_e *= d
                keep[e["offset"]:e["offset"] + n_e] = True
        mine = m._ps.flat[keep]
        w = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(w, mine)
        assert all(torch.equal(w[0], v) for v in w), f"{comm}: the replicas' weights diverged"
        if m._dist.peer is not None:
            assert not m._dist.peer.timed_out()
        finals[comm] = m.read_losses(eng)["loss"]
        say(f"{name} x{world}, {comm:12s} exchange: {float(t):.4f} ms/step, {B * world / float(t) * 1e3:.0f} images/s, "
            f"loss after 210 steps {finals[comm]:.4f}")
        del m, eng
    assert abs(finals["nccl"] - finals["peer"]) <= 2e-2 * abs(finals["nccl"]), finals
dist.destroy_process_group()
say("comm_check ok")
