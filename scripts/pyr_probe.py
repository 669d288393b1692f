"""Device time of the pyramid merge forward / adjoint / split at sub-problems of BASELINE configs[4] (which launch of the
level chain costs what): B = 128, C = 3, (H, levels) given on the command line as H:L pairs."""
import ctypes as Ct, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiscale_variational_autoencoder_b200 import _lib
from multiscale_variational_autoencoder_b200.engine import gaussian_kernel
lib = _lib.load()
dev = torch.device("cuda", 0)
B, C = int(os.environ.get("B", 128)), 3
s = torch.cuda.current_stream(dev).cuda_stream
taps = (Ct.c_float * 9)(*[float(v) for v in gaussian_kernel((3, 3), (2, 2)).astype(np.float32).ravel()])
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=8):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()                                  # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        e0.record(); _lib.check(fn()); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


for spec in sys.argv[1:]:
    H, L = map(int, spec.split(":"))
    W = H
    f = lambda i: torch.randn(B, H >> i, W >> i, C, device=dev)
    x = torch.rand(B, H, W, C, device=dev) * 255
    ys, dys, bands = [f(i) for i in range(L)], [f(i) for i in range(L)], [f(i) for i in range(L)]
    r0 = torch.empty(B, H, W, C, device=dev)
    P = lambda ts: (Ct.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    ws_s = torch.empty(lib.mvae_pyramid_split_workspace_bytes(B, H, W, C, L) // 4 + 1, device=dev)
    ws_m = torch.empty(lib.mvae_pyramid_merge_workspace_bytes(B, H, W, C, L) // 4 + 1, device=dev)
    yp, dyp, bp = P(ys), P(dys), P(bands)
    n = [B * (H >> i) * (W >> i) * C * 4 for i in range(L)]
    t_s = timed(lambda: lib.mvae_pyramid_split(x.data_ptr(), bp, ws_s.data_ptr(), B, H, W, C, L, 0.0, 255.0, taps, 3, 3, 0, s))
    t_m = timed(lambda: lib.mvae_pyramid_merge_fwd(yp, r0.data_ptr(), ws_m.data_ptr(), B, H, W, C, L, s))
    t_a = timed(lambda: lib.mvae_pyramid_merge_bwd(dys[0].data_ptr(), dyp, B, H, W, C, L, s))
    by = (n[0] + sum(n)) / 1e6
    print(f"H={H:4d} L={L}: split {t_s:7.1f} us ({by / t_s * 1e-3 * 1e3:6.0f} GB/s)   merge {t_m:7.1f} us ({by / t_m:6.0f} GB/s... MB/us)   "
          f"adjoint {t_a:7.1f} us ({sum(n) / 1e6 / t_a:6.2f} MB/us)   [{by:.1f} MB split/merge, {sum(n) / 1e6:.1f} MB adjoint]", flush=True)
