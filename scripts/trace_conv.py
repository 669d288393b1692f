"""Timeline of CTA 0 of the TMA conv kernel (globaltimer stamps per role): where does a tile's latency go?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)
B, H, W, Cc = 256, 16, 16, 32
names = {0: "setup done", 1: "TMA issued (tile*100+chunk)", 2: "chunk landed", 7: "chunk transformed", 8: "MMA warp saw chunk", 3: "MMAs issued", 9: "bias smem atomics done", 10: "transform warp done", 11: "producer done", 4: "accumulator ready", 5: "tile stored", 6: "CTA done"}
for (k, st, res) in [(1, 1, False), (3, 2, False)]:
    d = _lib.ConvDesc(B, H, W, Cc, k, k, st, st, Cc, 0, 1)
    Ho, Wo = -(-H // st), -(-W // st)
    x, y = f(B, H, W, Cc), f(B, Ho, Wo, Cc)
    w, bb = f(k, k, Cc, Cc) * .1, f(Cc)
    r = f(B, Ho, Wo, Cc) if res else None
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(lib.mvae_conv2d_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), bb.data_ptr(), 0, r.data_ptr() if res else 0, 1, y.data_ptr(), s))
    buf = torch.zeros(1 + 3000, dtype=torch.int64, device=dev)
    lib.mvae_debug_trace(buf.data_ptr())
    _lib.check(lib.mvae_conv2d_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), bb.data_ptr(), 0, r.data_ptr() if res else 0, 1, y.data_ptr(), s))
    torch.cuda.synchronize()
    lib.mvae_debug_trace(0)
    b = buf.cpu().tolist()
    n = b[0]
    ev = sorted([(b[3 + 3 * i], b[1 + 3 * i], b[2 + 3 * i]) for i in range(min(n, 1000))])
    t0 = ev[0][0]
    print(f"--- conv_fwd k{k} s{st} residual={res}: {n} events")
    for t, e, tile in ev:
        print(f"   {(t - t0) / 1e3:8.2f} us  tile {tile:4d}  {names[e]}")

# ---- wgrad timeline -----------------------------------------------------------------------------------------------
for (H_, W_) in [(4, 4), (16, 16)]:
    d = _lib.ConvDesc(B, H_, W_, Cc, 1, 1, 1, 1, Cc, 0, 1)
    x, dy = f(B, H_, W_, Cc), f(B, H_, W_, Cc)
    dw_, db_ = torch.zeros(Cc, Cc, device=dev), torch.zeros(Cc, device=dev)
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
    ev = sorted([(b[3 + 3 * i], b[1 + 3 * i], b[2 + 3 * i]) for i in range(min(n, 96))])
    t0 = ev[0][0]
    print(f"--- conv_wgrad 1x1 {H_}x{W_}: {n} events")
    for t, e, tile in ev:
        print(f"   {(t - t0) / 1e3:8.2f} us  {tile:4d}  {names[e]}")
