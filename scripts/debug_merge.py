import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
from oracle import mvae_oracle as O
lib = _lib.load()
S = lambda: torch.cuda.current_stream().cuda_stream
for shape, levels in [((2, 32, 32, 3), 2), ((2, 32, 32, 3), 3), ((2, 128, 128, 3), 4)]:
    B, H, W, Cc = shape
    ys = [torch.randn(B, H >> i, W >> i, Cc) for i in range(levels)]
    yd = [y.cuda() for y in ys]
    ptrs = (C.c_void_p * levels)(*[t.data_ptr() for t in yd])
    r0 = torch.empty(shape, device="cuda")
    ws = torch.empty(lib.mvae_pyramid_merge_workspace_bytes(B, H, W, Cc, levels) // 4 + 1, device="cuda")
    rc = lib.mvae_pyramid_merge_fwd(ptrs, r0.data_ptr(), ws.data_ptr(), B, H, W, Cc, levels, S())
    print(shape, levels, "rc", rc, _lib.last_error() if rc else "")
    torch.cuda.synchronize()
    ref = O.pyramid_merge_raw([y.double() for y in ys])
    print("  err", float((r0.double().cpu() - ref).abs().max()))
