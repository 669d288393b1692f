"""Steady-state device time of the conv_base kernels (3 -> 32 channels, 3x3) at the cfg2 level sizes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
from time_wide import timeit
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)
B = 256
for H in (32, 16):
    d = _lib.ConvDesc(B, H, H, 3, 3, 3, 1, 1, 32, 0, 1)
    x, y, dy = f(B, H, H, 3), f(B, H, H, 32), f(B, H, H, 32)
    w, bb = f(3, 3, 3, 32) * .1, f(32)
    dw_, db_ = torch.zeros_like(w), torch.zeros_like(bb)
    t = timeit("fwd", lambda s: lib.mvae_conv2d_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), bb.data_ptr(), 0, 0, 1, y.data_ptr(), s), n=20)
    print(f"conv_base {H}x{H}: fwd   {t:7.1f} us", flush=True)
    t = timeit("wgrad", lambda s: lib.mvae_conv2d_wgrad(C.byref(d), x.data_ptr(), 0, dy.data_ptr(), dw_.data_ptr(), db_.data_ptr(), s), n=20)
    print(f"conv_base {H}x{H}: wgrad {t:7.1f} us", flush=True)

# ---- decoder tail: BatchNorm statistics, BN + 1x1 -> 3 forward, backward (reduce + apply), colsum --------------------------
M, F, Co = 256 * 32 * 32, 32, 3
x, dyo, dx = f(M, F), f(M, Co), f(M, F)
sums = torch.zeros(2 * F, dtype=torch.float64, device=dev)
g_, be, mm, mv = f(F), f(F), torch.zeros(F, device=dev), torch.ones(F, device=dev)
w, b = f(F, Co) * .1, f(Co)
y, stats = f(M, Co), torch.zeros(2 * F, device=dev)
red = torch.zeros(F * Co + Co, device=dev)
G = [torch.zeros_like(t) for t in (g_, be, w, b)]
cs = torch.zeros(F, device=dev)
print(f"bn_stats        {timeit('bn_stats', lambda s: lib.mvae_bn_stats(x.data_ptr(), sums.data_ptr(), M, F, s), n=20):7.1f} us")
print(f"bn_convout_fwd  {timeit('fwd', lambda s: lib.mvae_bn_convout_fwd(x.data_ptr(), sums.data_ptr(), g_.data_ptr(), be.data_ptr(), mm.data_ptr(), mv.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), stats.data_ptr(), M, F, Co, 1e-4, 0.999, 1, s), n=20):7.1f} us")
print(f"bn_convout_bwd  {timeit('bwd', lambda s: lib.mvae_bn_convout_bwd(x.data_ptr(), dyo.data_ptr(), stats.data_ptr(), g_.data_ptr(), be.data_ptr(), w.data_ptr(), red.data_ptr(), dx.data_ptr(), G[0].data_ptr(), G[1].data_ptr(), G[2].data_ptr(), G[3].data_ptr(), M, F, Co, s), n=20):7.1f} us  (reduce + apply)")
print(f"colsum          {timeit('colsum', lambda s: lib.mvae_colsum(x.data_ptr(), cs.data_ptr(), M, F, s), n=20):7.1f} us")
