"""One GPU: does the stream structure of the early gradient exchange (a kernel behind each big Dense weight gradient on its
side stream) cost time by itself?  The exchange kernels are replaced by a 4-byte memset.  (A side stream that joined every
level's decoder half did: 1.34 -> 1.53 ms per cfg2 step with nothing but the joins.)"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multiscale_variational_autoencoder_b200 import MultiscaleVAE, _lib

dev = torch.device("cuda", 0)
cfg, B, _ = bench.CONFIGS["cfg2"]
g = torch.Generator().manual_seed(1)
x = torch.rand(B, *cfg["input_dims"], generator=g) * 255
eps = [torch.randn(B, z, generator=g) for z in cfg["z_dims"]]
lib = _lib.load()
for mode in (sys.argv[1:] or ("plain", "fake-late", "fake-early", "fake-early-kernel")):
    m = MultiscaleVAE(**cfg, precision="tf32", device=dev, seed=7)
    m.compile(0.01, 1.0, 0.1)
    if mode != "plain":
        scratch = torch.zeros(16, device=dev)
        peer = types.SimpleNamespace(
            allreduce=(lambda stream=None, **kw: _lib.check(lib.mvae_accumulate(
                scratch.data_ptr(), scratch.data_ptr() + 16, 1, 1.0, stream or torch.cuda.current_stream(dev).cuda_stream)))
            if mode.endswith("kernel") else
            (lambda stream=None, **kw: _lib.check(lib.mvae_memset_zero(
                scratch.data_ptr(), 4, stream or torch.cuda.current_stream(dev).cuda_stream))),
            timed_out=lambda: False, CHANNELS=16)
        from multiscale_variational_autoencoder_b200.dist import GradAllReduce
        ar = GradAllReduce.__new__(GradAllReduce)
        ar.ps, ar.peer, ar.world, ar.rank = m._ps, peer, 1, 0
        m._dist = ar
        m._dp_ingraph = False
        m._dp_early = ar.early_ranges() if mode.startswith("fake-early") else {}
    eng = m._engine(B, True)
    m._load_input(eng, x.numpy())
    m._load_eps(eng, eps)
    for _ in range(10):
        m.train_step_device(eng)
    torch.cuda.synchronize()
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(300):
        m.train_step_device(eng)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    # host cost of the two graph launches alone, GPU idle in between (each step synchronised)
    th = 0.0
    for _ in range(50):
        torch.cuda.synchronize()
        a = time.perf_counter()
        m.train_step_device(eng)
        th += time.perf_counter() - a
    print(f"{mode:18s} {e0.elapsed_time(e1) / 300:.4f} ms/step  ({eng.kernels_per_step} kernel nodes)  host enqueue loop "
          f"{(t1 - t0) / 300 * 1e3:.4f} ms/step, with final sync {(t2 - t0) / 300 * 1e3:.4f}; launch call on an idle GPU {th / 50 * 1e3:.4f} ms", flush=True)
