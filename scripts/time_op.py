"""Steady-state device time of single C-ABI calls: N back-to-back launches captured in a CUDA graph, replayed between events."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)

def timeit(name, fn, n=50):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        s = st.cuda_stream
        for _ in range(3): _lib.check(fn(s), name)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): _lib.check(fn(torch.cuda.current_stream().cuda_stream), name)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); st.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / n * 1e3:8.2f} us / launch")

B, Cc, HW = 256, 32, 256
gap, gate = f(B, Cc).abs(), f(B, Cc)
w0, w1, b0, b1, gam, bet = f(Cc, Cc) * .1, f(Cc, Cc) * .1, f(Cc), f(Cc), f(Cc), f(Cc)
mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
ws = torch.empty(lib.mvae_se_gate_ws_floats(B, Cc), device=dev)
timeit("se_gate_fwd B256 C32", lambda s: lib.mvae_se_gate_fwd(gap.data_ptr(), w0.data_ptr(), b0.data_ptr(), gam.data_ptr(), bet.data_ptr(), w1.data_ptr(), b1.data_ptr(), mm.data_ptr(), mv.data_ptr(), gate.data_ptr(), ws.data_ptr(), B, Cc, HW, 1e-3, 0.99, 1, s))
dg, dgap = f(B, Cc), f(B, Cc)
G = [torch.zeros_like(t) for t in (w0, b0, gam, bet, w1, b1)]
timeit("se_gate_bwd B256 C32", lambda s: lib.mvae_se_gate_bwd(dg.data_ptr(), w0.data_ptr(), gam.data_ptr(), bet.data_ptr(), w1.data_ptr(), ws.data_ptr(), dgap.data_ptr(), *[t.data_ptr() for t in G], B, Cc, HW, s))
for (H, W) in [(16, 16), (32, 32), (4, 4)]:
    a, u, dv, da = f(B, H, W, Cc), f(B, H, W, Cc), f(B, H, W, Cc), f(B, H, W, Cc)
    wd, bd = f(3, 3, Cc), f(Cc)
    gs = torch.zeros(B, Cc, device=dev)
    dwd, dbd = torch.zeros_like(wd), torch.zeros_like(bd)
    timeit(f"dw_fwd {H}x{W}", lambda s: lib.mvae_dwconv3x3_fwd(a.data_ptr(), wd.data_ptr(), bd.data_ptr(), u.data_ptr(), gs.data_ptr(), B, H, W, Cc, s))
    timeit(f"dw_bwd {H}x{W}", lambda s: lib.mvae_dwconv3x3_bwd(a.data_ptr(), u.data_ptr(), dv.data_ptr(), gate.data_ptr(), dgap.data_ptr(), wd.data_ptr(), da.data_ptr(), dwd.data_ptr(), dbd.data_ptr(), B, H, W, Cc, s))
    timeit(f"dgate_reduce {H}x{W}", lambda s: lib.mvae_se_dgate_reduce(dv.data_ptr(), u.data_ptr(), gs.data_ptr(), B, H * W, Cc, s))
    for (k, st_) in [(1, 1), (3, 2)]:
        d = _lib.ConvDesc(B, H, W, Cc, k, k, st_, st_, Cc, 0, 1)
        Ho, Wo = -(-H // st_), -(-W // st_)
        x, y, dy, dx = f(B, H, W, Cc), f(B, Ho, Wo, Cc), f(B, Ho, Wo, Cc), f(B, H, W, Cc)
        w, bb = f(k, k, Cc, Cc) * .1, f(Cc)
        dw_, db_ = torch.zeros_like(w), torch.zeros_like(bb)
        timeit(f"conv_fwd {H}x{W} k{k}s{st_}", lambda s: lib.mvae_conv2d_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), bb.data_ptr(), 0, 0, 1, y.data_ptr(), s))
        timeit(f"conv_dgrad {H}x{W} k{k}s{st_}", lambda s: lib.mvae_conv2d_dgrad(C.byref(d), dy.data_ptr(), w.data_ptr(), 0, 0, 0, 0, dx.data_ptr(), s))
        timeit(f"conv_wgrad {H}x{W} k{k}s{st_}", lambda s: lib.mvae_conv2d_wgrad(C.byref(d), x.data_ptr(), 0, dy.data_ptr(), dw_.data_ptr(), db_.data_ptr(), s))
