"""Timeline of CTA 0 of the TMA wgrad kernel for a 3x3 convolution."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)
names = {0: "setup done", 1: "TMA issued", 2: "chunk landed", 7: "chunk transformed", 8: "MMA warp saw chunk", 3: "MMAs issued", 9: "bias done", 10: "transform warp done", 11: "producer done", 4: "accumulator ready", 5: "tile stored", 6: "CTA done"}
B, H, Cc, k, st = 256, 16, 32, 3, 1
d = _lib.ConvDesc(B, H, H, Cc, k, k, st, st, Cc, 0, 1)
x, dy = f(B, H, H, Cc), f(B, H, H, Cc)
dw_, db_ = torch.zeros(k, k, Cc, Cc, device=dev), torch.zeros(Cc, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.mvae_conv2d_wgrad(C.byref(d), x.data_ptr(), 0, dy.data_ptr(), dw_.data_ptr(), db_.data_ptr(), s))
buf = torch.zeros(1 + 3000, dtype=torch.int64, device=dev)
lib.mvae_debug_trace(buf.data_ptr())
_lib.check(lib.mvae_conv2d_wgrad(C.byref(d), x.data_ptr(), 0, dy.data_ptr(), dw_.data_ptr(), db_.data_ptr(), s))
torch.cuda.synchronize()
lib.mvae_debug_trace(0)
b = buf.cpu().tolist()
n = b[0]
ev = sorted([(b[3 + 3 * i], b[1 + 3 * i], b[2 + 3 * i]) for i in range(min(n, 900))])
t0 = ev[0][0]
for t, e, tile in ev[:70] + ev[-14:]:
    print(f"   {(t - t0) / 1e3:8.2f} us  {tile:4d}  {names.get(e, e)}")
