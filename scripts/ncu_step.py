"""Two eager (no CUDA graph, one stream) training steps of a bench config: the target of the ncu launch-list pass.
   ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of step 1> -c <launches of step 2> ... """
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multiscale_variational_autoencoder_b200 import MultiscaleVAE

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--precision", default="tf32")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
cfg, B, _ = bench.CONFIGS[a.config]
m = MultiscaleVAE(**cfg, precision=a.precision)
m.compile(0.01, 1.0, 0.1)
m.use_cuda_graph = m.parallel_levels = False
eng = m._engine(B, True)
eng.defer_serial = True        # weight gradients as the batched launches of the captured graph
g = torch.Generator().manual_seed(0)
eng.x.copy_(torch.rand(B, *cfg["input_dims"], generator=g) * 255)
for e in eng.eps:
    e.normal_(0, 1)
for _ in range(a.steps):
    m.train_step_device(eng)
torch.cuda.synchronize()
print("loss", m.read_losses(eng))
