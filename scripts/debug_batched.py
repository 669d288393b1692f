import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_step as S
cfg = dict(input_dims=(32, 32, 3), z_dims=[16, 8], sample_std=0.5,
           encoder={"filters": [32, 32], "kernel_size": [(3, 3)] * 2, "strides": [(2, 2), (1, 1)]})
def run(graph, batch, steps=3, prec="tf32"):
    m, _, x, eps = S.make_pair(cfg, 8, seed=5, precision=prec)
    w0 = m.state_dict()
    m.compile(0.01, 1.0, 0.1)
    m.use_cuda_graph = graph; m.parallel_levels = graph
    m._engine(8, True).batch_levels = batch
    for _ in range(steps):
        m.train_on_batch(x.numpy(), eps)
    return w0, m.state_dict()
def cmp(a, b, w0):
    upd = max(float((a[k] - w0[k]).abs().max()) for k in a)
    worst = max((float((a[k] - b[k]).abs().max()) / upd, k) for k in a)
    return worst
for prec in ("tf32", "fp32"):
    w0, e1 = run(False, False, prec=prec)
    _, e2 = run(False, False, prec=prec)
    _, g1 = run(True, False, prec=prec)
    _, g2 = run(True, True, prec=prec)
    print(prec, "eager vs eager      ", cmp(e1, e2, w0))
    print(prec, "eager vs graph      ", cmp(e1, g1, w0))
    print(prec, "eager vs graph+batch", cmp(e1, g2, w0))
    print(prec, "graph vs graph+batch", cmp(g1, g2, w0))
    for steps in (1,):
        w0, a = run(False, False, steps, prec); _, b = run(True, True, steps, prec)
        print(prec, "1 step eager vs graph+batch", cmp(a, b, w0))
