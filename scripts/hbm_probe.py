"""Why does the same reconstruction-loss launch take 119 us in one run and 213 us in another?
Times the cfg5 kernels back to back, behind a GPU spin, after host-side idle time, and prints every sample."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
B, H, W, Cc = 128, 512, 512, 3
f32 = dict(dtype=torch.float32, device=dev)
x = torch.rand(B, H, W, Cc, **f32) * 255
r0 = torch.randn(B, H, W, Cc, **f32) * 0.5
dr0 = torch.empty(B, H, W, Cc, **f32)
sums = torch.zeros(B * 7, **f32)
s = torch.cuda.current_stream().cuda_stream
fwd = lambda: lib.mvae_recon_loss_fwd(r0.data_ptr(), x.data_ptr(), 0, sums.data_ptr(), B, H, W, Cc, 0.0, 255.0, s)
bwd = lambda: lib.mvae_recon_loss_bwd(r0.data_ptr(), x.data_ptr(), sums.data_ptr(), dr0.data_ptr(), B, H, W, Cc, 0.0, 255.0, 1.0 / B, s)
copy = lambda: dr0.copy_(r0)


def sample(fn, n, spin, idle=0.0):
    out = []
    for _ in range(n):
        if idle:
            torch.cuda.synchronize(); time.sleep(idle)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if spin:
            torch.cuda._sleep(spin)
        e0.record(); fn(); e1.record()
        out.append((e0, e1))
    torch.cuda.synchronize()
    return [round(a.elapsed_time(b) * 1e3, 1) for a, b in out]


def clocks():
    import subprocess
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader"],
                          capture_output=True, text=True).stdout.strip()


for name, fn, nbytes in (("recon_fwd", fwd, 8.0 * x.numel()), ("recon_bwd", bwd, 12.0 * x.numel()), ("torch copy", copy, 8.0 * x.numel())):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(f"== {name}: {nbytes / 1e6:.0f} MB   clocks {clocks()}")
    for label, kw in (("back to back", dict(spin=0)), ("behind a 0.2 ms spin", dict(spin=400000)),
                      ("behind a 5 ms spin", dict(spin=10000000)), ("after 50 ms host idle", dict(spin=0, idle=0.05)),
                      ("after 0.5 s host idle", dict(spin=0, idle=0.5)), ("back to back again", dict(spin=0))):
        t = sample(fn, 6 if kw.get("idle", 0) >= 0.5 else 10, **kw)
        print(f"   {label:24s} {t}   best {nbytes / min(t) / 1e3:.0f} GB/s  median {nbytes / sorted(t)[len(t) // 2] / 1e3:.0f} GB/s   clocks {clocks()}")
