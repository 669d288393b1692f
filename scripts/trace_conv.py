"""Timeline of CTA 0 of the TMA conv kernel (globaltimer stamps per role): where does a tile's latency go?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)
B, H, W, Cc = 256, 16, 16, 32
names = {0: "setup done", 1: "TMA issued (tile*100+chunk)", 2: "chunk landed", 7: "chunk transformed", 8: "MMA warp saw chunk", 3: "MMAs issued", 9: "bias smem atomics done", 10: "transform warp done", 11: "producer done", 4: "accumulator ready", 5: "tile stored", 6: "CTA done", 12: "epi: tmem loaded", 13: "epi: staging buffer free", 14: "epi: staged", 15: "epi: fenced + synced"}
for (k, st, res) in [(1, 1, False), (1, 1, True)]:
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

# ---- squeeze-excite gate forward timeline ------------------------------------------------------------------------------
sen = {0: "start", 1: "weights + gap staged", 2: "dense0 done", 3: "local stats done", 4: "cluster barrier passed", 5: "batch stats combined", 6: "dense1 + gate stored", 7: "final cluster wait done"}
gap, gate = f(B, Cc).abs(), f(B, Cc)
w0, w1, b0, b1, gam, bet = f(Cc, Cc) * .1, f(Cc, Cc) * .1, f(Cc), f(Cc), f(Cc), f(Cc)
mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
ws = torch.empty(lib.mvae_se_gate_ws_floats(B, Cc), device=dev)
call = lambda s: lib.mvae_se_gate_fwd(gap.data_ptr(), w0.data_ptr(), b0.data_ptr(), gam.data_ptr(), bet.data_ptr(), w1.data_ptr(), b1.data_ptr(), mm.data_ptr(), mv.data_ptr(), gate.data_ptr(), ws.data_ptr(), B, Cc, 256, 1e-3, 0.99, 1, s)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3): _lib.check(call(s))
buf = torch.zeros(1 + 3000, dtype=torch.int64, device=dev)
lib.mvae_debug_trace(buf.data_ptr())
_lib.check(call(s)); torch.cuda.synchronize()
lib.mvae_debug_trace(0)
b = buf.cpu().tolist()
ev = sorted([(b[3 + 3 * i], b[1 + 3 * i]) for i in range(b[0])])
print("--- se_gate_fwd")
for t, e in ev:
    print(f"   {(t - ev[0][0]) / 1e3:8.2f} us  {sen[e]}")
