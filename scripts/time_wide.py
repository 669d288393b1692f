"""Steady-state device time of the wide convolutions of cfg4 (B = 64): graph-replayed back-to-back launches."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
if os.environ.get("ALT_LIB"):            # A/B against another build of the library
    _lib.LIB_PATH = os.environ["ALT_LIB"]
    _lib._needs_build = lambda: False
lib = _lib.load()
dev = torch.device("cuda", 0)
f = lambda *s: torch.randn(*s, device=dev)


def timeit(name, fn, n=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2): _lib.check(fn(st.cuda_stream), name)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): _lib.check(fn(torch.cuda.current_stream().cuda_stream), name)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); st.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = 64
for (H, Cin, Cout, k, s_) in [] if __name__ != "__main__" else [(128, 64, 128, 3, 1), (256, 64, 128, 3, 2), (128, 128, 128, 1, 1), (256, 64, 64, 1, 1)]:
    d = _lib.ConvDesc(B, H, H, Cin, k, k, s_, s_, Cout, 0, 1)
    Ho = -(-H // s_)
    x, y, dy, dx = f(B, H, H, Cin), f(B, Ho, Ho, Cout), f(B, Ho, Ho, Cout), f(B, H, H, Cin)
    w, bb = f(k, k, Cin, Cout) * .05, f(Cout)
    dw_, db_ = torch.zeros_like(w), torch.zeros_like(bb)
    gb = (x.numel() + y.numel()) * 4 / 1e3
    fl = 2.0 * B * Ho * Ho * Cout * Cin * k * k / 1e6
    t = timeit("fwd", lambda s: lib.mvae_conv2d_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), bb.data_ptr(), 0, 0, 1, y.data_ptr(), s))
    print(f"{H}x{H}x{Cin}->{Cout} k{k}s{s_}: fwd   {t:8.1f} us  {gb / t:6.0f} GB/s  {fl / t:6.1f} TF/s", flush=True)
    t = timeit("dgrad", lambda s: lib.mvae_conv2d_dgrad(C.byref(d), dy.data_ptr(), w.data_ptr(), 0, 0, 0, 0, dx.data_ptr(), s))
    print(f"{H}x{H}x{Cin}->{Cout} k{k}s{s_}: dgrad {t:8.1f} us  {gb / t:6.0f} GB/s  {fl / t:6.1f} TF/s", flush=True)
    lib.mvae_set_wgrad_sm_share(148)
    t = timeit("wgrad", lambda s: lib.mvae_conv2d_wgrad(C.byref(d), x.data_ptr(), 0, dy.data_ptr(), dw_.data_ptr(), db_.data_ptr(), s))
    print(f"{H}x{H}x{Cin}->{Cout} k{k}s{s_}: wgrad {t:8.1f} us  {gb / t:6.0f} GB/s  {fl / t:6.1f} TF/s  (all SMs)", flush=True)
    del x, y, dy, dx
