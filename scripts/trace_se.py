"""Phase timeline (globaltimer) of CTA 0 of the squeeze-excite gate kernels."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)
B, Cc = 256, 32
gap, gate = f(B, Cc).abs(), f(B, Cc)
w0, w1, b0, b1, gam, bet = f(Cc, Cc) * .1, f(Cc, Cc) * .1, f(Cc), f(Cc), f(Cc), f(Cc)
mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
ws = torch.empty(lib.mvae_se_gate_ws_floats(B, Cc), device=dev)
dg, dgap = f(B, Cc), f(B, Cc)
G = [torch.zeros_like(t) for t in (w0, b0, gam, bet, w1, b1)]
fwd = lambda s: lib.mvae_se_gate_fwd(gap.data_ptr(), w0.data_ptr(), b0.data_ptr(), gam.data_ptr(), bet.data_ptr(), w1.data_ptr(), b1.data_ptr(), mm.data_ptr(), mv.data_ptr(), gate.data_ptr(), ws.data_ptr(), B, Cc, 256, 1e-3, 0.99, 1, s)
bwd = lambda s: lib.mvae_se_gate_bwd(dg.data_ptr(), w0.data_ptr(), gam.data_ptr(), bet.data_ptr(), w1.data_ptr(), ws.data_ptr(), dgap.data_ptr(), *[t.data_ptr() for t in G], B, Cc, 256, s)
s = torch.cuda.current_stream().cuda_stream
for name, call in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): _lib.check(call(s))
    for rep in range(2):
        buf = torch.zeros(1 + 3000, dtype=torch.int64, device=dev)
        lib.mvae_debug_trace(buf.data_ptr())
        _lib.check(call(s)); torch.cuda.synchronize()
        lib.mvae_debug_trace(0)
        b = buf.cpu().tolist()
        ev = sorted([(b[3 + 3 * i], b[1 + 3 * i]) for i in range(b[0])])
        print(f"--- se_gate_{name}: " + "  ".join(f"{e}:{(t - ev[0][0]) / 1e3:.2f}" for t, e in ev))
