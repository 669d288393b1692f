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
