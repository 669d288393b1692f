"""How much of the cfg2 step is level 0's dependency chain?  Times the graph-replayed training step of the cfg2 geometry
truncated to 1..5 pyramid levels (level 0 is identical in all of them), and the serial-stream variant of the full model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import MultiscaleVAE

dev = torch.device("cuda", 0)
Z = [128, 64, 32, 16, 8]
B = int(os.environ.get("B", "256"))


def run(levels, serial=False, steps=30):
    m = MultiscaleVAE(input_dims=(32, 32, 3), z_dims=Z[:levels], sample_std=0.5, precision="tf32", device=dev,
                      encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (1, 1), (1, 1)]})
    m.compile(0.01, 1.0, 0.1)
    if serial:
        m.parallel_levels = False
    eng = m._engine(B, True)
    eng.x.copy_(torch.rand(B, 32, 32, 3) * 255)
    for e in eng.eps:
        e.normal_()
    for _ in range(5):
        m.train_step_device(eng)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.train_step_device(eng)
    e1.record()
    torch.cuda.synchronize()
    print(f"levels={levels} serial={serial}: {e0.elapsed_time(e1) / steps:.3f} ms/step", flush=True)


for L in (2, 3, 4, 5):
    run(L)
run(5, serial=True)
