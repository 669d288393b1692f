"""Training steps through the CUDA-graph path (level-batched launches, side streams): the target of ncu launch lists that
should show the kernels exactly as the timed bench runs them."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multiscale_variational_autoencoder_b200 import MultiscaleVAE

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--precision", default="tf32")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
cfg, B, _ = bench.CONFIGS[a.config]
m = MultiscaleVAE(**cfg, precision=a.precision)
m.compile(0.01, 1.0, 0.1)
eng = m._engine(B, True)
g = torch.Generator().manual_seed(0)
eng.x.copy_(torch.rand(B, *cfg["input_dims"], generator=g) * 255)
for e in eng.eps:
    e.normal_(0, 1)
for _ in range(a.steps + 1):        # the first call warms up eagerly and captures; the others replay
    m.train_step_device(eng)
torch.cuda.synchronize()
print("loss", m.read_losses(eng))
