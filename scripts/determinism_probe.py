"""Run-to-run reproducibility of three fp32 training steps: per-parameter max difference against the first run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import MultiscaleVAE
cfg = dict(input_dims=(32, 32, 3), z_dims=[128, 64, 32], sample_std=0.5,
           encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (2, 2), (1, 1)]})
B = 8
g = torch.Generator().manual_seed(5)
xs = [(torch.rand(B, 32, 32, 3, generator=g) * (80.0 * (i + 1))) for i in range(3)]
es = [[torch.randn(B, z, generator=g) for z in cfg["z_dims"]] for _ in range(3)]
graph = os.environ.get("GRAPH", "1") == "1"
ref = None
for rep in range(6):
    m = MultiscaleVAE(**cfg, seed=3)
    m.compile(0.01, 1.0, 0.1)
    m.use_cuda_graph = m.parallel_levels = graph
    eng = m._engine(B, True)
    gr = []
    for i in range(3):
        m._load_input(eng, xs[i].numpy())
        m._load_eps(eng, es[i])
        m.train_step_device(eng)
        torch.cuda.synchronize()
        gr.append({k: v.clone() for k, v in m._ps.state_dict(grads=True).items()})
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    if ref is None:
        ref = (sd, gr)
        continue
    worst = sorted(((float((sd[k] - ref[0][k]).abs().max()) / max(float(ref[0][k].abs().max()), 1e-2), k) for k in sd), reverse=True)[:3]
    gw = []
    for i in range(3):
        gw.append(max((float((gr[i][k] - ref[1][i][k]).abs().max()) / max(float(ref[1][i][k].abs().max()), 1e-6), k) for k in gr[i]))
    print(f"rep {rep}: params {[(round(a, 6), k) for a, k in worst]}  grads/step {[(round(a, 5), k) for a, k in gw]}", flush=True)
