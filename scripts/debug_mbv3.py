import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from multiscale_variational_autoencoder_b200 import engine as E
import test_gpu_kernels as K

def run(prec, B=8, H=16, W=16, Cc=32, F=32):
    ps = E.ParamStore(torch.device("cuda", 0), seed=3)
    E.declare_mbv3(ps, "m_", Cc, F)
    ps.finalize()
    g = torch.Generator().manual_seed(8)
    for n in ps.entries:
        if n.endswith("bias") or n.endswith("beta"):
            ps.view(n).copy_(torch.randn(ps.view(n).shape, generator=g) * 0.1)
    eng = K.MiniTrain(ps, B, prec)
    x = K.rnd((B, H, W, Cc), 1).cuda()
    gy = K.rnd((B, H, W, Cc), 2).cuda()
    xt = E.T(x, torch.empty_like(x))
    op = E.MobileNetV3(eng, xt, "m_", F)
    op.fwd()
    op.y.grad.copy_(gy)
    op.bwd()
    torch.cuda.synchronize()
    out = dict(a=op.a, u=op.u, gate=op.gate, y=op.y.data, dv=op.dv, dg=torch.zeros(1), dgap=op.dgap, da=op.da, dx=xt.grad)
    out.update({"g:" + k: v.cuda() for k, v in ps.state_dict(grads=True).items()})
    return {k: v.clone() for k, v in out.items()}

for shape in [(8, 16, 16, 32, 32), (4, 16, 16, 64, 128)]:
    a, b = run(0, *shape), run(1, *shape)
    print("shape", shape)
    for k in a:
        d = (a[k] - b[k]).double()
        print(f"  {k:45s} max-rel {float(d.abs().max() / max(float(a[k].abs().max()), 1e-12)):.2e}  l2-rel {float(d.norm() / max(float(a[k].double().norm()), 1e-12)):.2e}")
