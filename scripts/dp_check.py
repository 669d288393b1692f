"""Data-parallel check on N GPUs (torchrun): (1) the gradients every rank holds after the in-graph per-level exchange equal
the mean over ranks of the per-replica gradients (computed locally for every rank's shard with the same weights: BatchNorm
uses per-replica statistics, as Keras does under MirroredStrategy); (2) step time with the per-level in-graph exchange
against one all-reduce between the two graphs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dp_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from multiscale_variational_autoencoder_b200 import MultiscaleVAE

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, B, _ = bench.CONFIGS[name]
B = int(os.environ.get("B", B))
H, W, C = cfg["input_dims"]


def data(r):
    g = torch.Generator().manual_seed(1000 + r)
    return torch.rand(B, H, W, C, generator=g) * 255, [torch.randn(B, z, generator=g) for z in cfg["z_dims"]]


def make(ingraph):
    os.environ["MVAE_DP_INGRAPH"] = "1" if ingraph else "0"
    m = MultiscaleVAE(**cfg, precision=os.environ.get("PREC", "fp32"), device=dev, seed=7)
    m.compile(0.01, 1.0, 0.1)
    m.enable_data_parallel()
    return m


# ---- (1) gradient equality (fp32 kernels: compared at 1e-3 of the largest gradient: a wrong range or a missing exchange shows as O(1)) -------------------------------------
m = make(os.environ.get("DP_CHECK_INGRAPH") == "1")
eng = m._engine(B, True)
x, eps = data(rank)
m._load_input(eng, x.numpy())
m._load_eps(eng, eps)
w0 = m._ps.flat.clone()
m._step_body(eng)                                   # eager: forward, backward (+ per-level exchange when in-graph)
if not m._dp_ingraph and m._dist.peer is None:
    m._dist.allreduce()                             # NCCL transport: the exchange follows the step (the peer kernel is inside it)
torch.cuda.synchronize()
got = m._ps.grads.clone() / world
ref = torch.zeros_like(got)
solo = MultiscaleVAE(**cfg, precision=os.environ.get("PREC", "fp32"), device=dev, seed=7)
solo.compile(0.01, 1.0, 0.1)
solo._ps.flat.copy_(w0)
e2 = solo._engine(B, True)
for r in range(world):
    xr, er = data(r)
    solo._load_input(e2, xr.numpy())
    solo._load_eps(e2, er)
    e2.forward_backward(parallel=False)
    torch.cuda.synchronize()
    ref += solo._ps.grads / world
scale = float(ref.abs().max())
err = float((got - ref).abs().max()) / scale
print(f"[rank {rank}] exchanged gradients vs mean of per-replica gradients: max err {err:.2e} of max |g| {scale:.3e}", flush=True)
assert err <= 1e-3, err          # fp32 kernels: atomic summation order differs between the runs
del solo, e2

# ---- (2) step time ---------------------------------------------------------------------------------------------------
for ingraph in ((False, True) if os.environ.get("DP_CHECK_INGRAPH") == "1" else (False,)):
    os.environ["PREC"] = "tf32"
    m = make(ingraph)
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
    if rank == 0:
        print(f"{name} x{world}: {'per-level exchange inside the graph' if ingraph else 'one all-reduce between the graphs'}: "
              f"{float(t):.4f} ms/step, {B * world / float(t) * 1e3:.0f} images/s", flush=True)
    del m, eng
dist.destroy_process_group()
