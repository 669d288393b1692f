"""Structured probes of the tcgen05 conv kernel (debug aid): 1x1 conv == GEMM y = x W."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
from multiscale_variational_autoencoder_b200._lib import ConvDesc
lib = _lib.load()
S = lambda: torch.cuda.current_stream().cuda_stream
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)

def run(x, w, mode="fwd", k=1, s=1, prec=1):
    B, H, W_, Cin = x.shape
    if mode == "fwd":
        Cout = w.shape[-1]
        d = ConvDesc(B, H, W_, Cin, k, k, s, s, Cout, 0, prec)
        y = torch.full((B, -(-H // s), -(-W_ // s), Cout), -7.0, device="cuda")
        _lib.check(lib.mvae_conv2d_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), 0, 0, 0, 0, y.data_ptr(), S()))
    else:   # dgrad: x plays dy (B,Ho,Wo,Cout); w (k,k,Cin,Cout)
        Cin_f, Cout = w.shape[2], w.shape[3]
        Hf, Wf = H * s, W_ * s
        d = ConvDesc(B, Hf, Wf, Cin_f, k, k, s, s, Cout, 0, prec)
        y = torch.full((B, Hf, Wf, Cin_f), -7.0, device="cuda")
        _lib.check(lib.mvae_conv2d_dgrad(C.byref(d), x.data_ptr(), w.data_ptr(), 0, 0, 0, 0, y.data_ptr(), S()))
    torch.cuda.synchronize()
    return y

for mode in ("fwd", "dgrad"):
    print("=====", mode)
    B, H, W_, Cc = 2, 16, 16, 32
    x = torch.randn(B, H, W_, Cc, device="cuda")
    w = torch.eye(Cc, device="cuda").view(1, 1, Cc, Cc).contiguous()
    y = run(x, w, mode)
    print("identity W: max|y-x| =", float((y - x).abs().max()), " y[0,0,0,:8]", y[0, 0, 0, :8].tolist(), " x", x[0, 0, 0, :8].tolist())
    # W with one nonzero row / col
    w = torch.zeros(1, 1, Cc, Cc, device="cuda")
    w[0, 0, 3, :] = torch.arange(1, Cc + 1, device="cuda").float()      # ci=3 -> co: co+1
    x = torch.zeros(B, H, W_, Cc, device="cuda")
    x[..., 3] = 1.0
    x[..., 5] = 100.0
    y = run(x, w, mode)
    print("W[3,:]=1..32 ; x[...,3]=1,x[...,5]=100: y[0,0,0,:] =", y[0, 0, 0, :].tolist())
    print("   y[1,7,9,:8] =", y[1, 7, 9, :8].tolist())
    # row id probe
    x = torch.zeros(B, H, W_, Cc, device="cuda")
    x.view(-1, Cc)[:, 0] = torch.arange(B * H * W_, device="cuda").float()
    w = torch.zeros(1, 1, Cc, Cc, device="cuda")
    w[0, 0, 0, :] = 1.0
    y = run(x, w, mode).view(-1, Cc)
    print("row-id probe: y[:6,0] =", y[:6, 0].tolist(), " y[126:131,0] =", y[126:131, 0].tolist(), " y[300,:4]", y[300, :4].tolist())
    # random full check vs torch
    x = torch.randn(B, H, W_, Cc, device="cuda")
    w = torch.randn(1, 1, Cc, Cc, device="cuda")
    y = run(x, w, mode)
    ref = x.view(-1, Cc) @ (w.view(Cc, Cc) if mode == "fwd" else w.view(Cc, Cc).t())
    print("random 1x1 relerr", float((y.view(-1, Cc) - ref).abs().max() / ref.abs().max()))
    y0 = run(x, w, mode, prec=0)
    print("fp32 path relerr", float((y0.view(-1, Cc) - ref).abs().max() / ref.abs().max()))
print("tc launches", lib.mvae_tc_launch_count())
