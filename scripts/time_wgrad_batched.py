"""Steady-state time of mvae_conv2d_wgrad_batched for n = 1..8 identical 1x1 problems (CUDA-graph replay)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
from multiscale_variational_autoencoder_b200.engine import _pa, _descs
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)
B, H, W, Cc = 256, 16, 16, 32


def timeit(name, fn, n=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): _lib.check(fn(st.cuda_stream), name)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): _lib.check(fn(torch.cuda.current_stream().cuda_stream), name)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); st.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / n * 1e3:8.2f} us / launch", flush=True)


for k, s_ in ((1, 1), (3, 2), (3, 1)):
    for n in (1, 8):
        Hi = H * s_
        d = [_lib.ConvDesc(B, Hi, Hi, Cc, k, k, s_, s_, Cc, 0, 1) for _ in range(n)]
        xs, dys = [f(B, Hi, Hi, Cc) for _ in range(n)], [f(B, H, W, Cc) for _ in range(n)]
        dws, dbs = [torch.zeros(k, k, Cc, Cc, device=dev) for _ in range(n)], [torch.zeros(Cc, device=dev) for _ in range(n)]
        P = lambda ts: _pa([t.data_ptr() for t in ts])
        da = _descs(d)
        timeit(f"wgrad_batched k{k}s{s_} n={n}", lambda s: lib.mvae_conv2d_wgrad_batched(n, da, P(xs), None, P(dys), P(dws), P(dbs), s))
