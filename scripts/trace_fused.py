"""Timeline of CTA 0 of the fused mobilenetV3 forward launches (globaltimer stamps): where does a launch's latency go?
Runs a chain of three blocks (F1, F2F1, F2F1, F2) back to back on one stream, eager and inside a CUDA graph."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiscale_variational_autoencoder_b200 import MultiscaleVAE, _lib
import bench
lib = _lib.load()
dev = torch.device("cuda", 0)
names = {0: "kernel start", 1: "setup done (barriers, TMEM)", 2: "weights staged", 3: "batch stats done", 4: "gate of tile done",
         5: "tile landed", 6: "operand rounded", 7: "conv2 MMA done", 8: "y written, store issued", 9: "conv0 MMA done",
         10: "a written", 11: "depthwise done, u store issued", 12: "image sums / h done", 13: "teardown done"}
cfg, B, _ = bench.CONFIGS["cfg2"]
B = int(os.environ.get("B", B))
m = MultiscaleVAE(**cfg, precision="tf32", device=dev)
m.compile(0.01, 1.0, 0.1)
eng = m._engine(B, True)
eng.x.copy_(torch.rand(B, 32, 32, 3) * 255)
for e in eng.eps:
    e.normal_()
bnames = {0: "kernel start", 1: "setup done", 2: "weights staged", 3: "BN-bwd sums done", 4: "dgap of tile done", 5: "tile landed",
          6: "dy rounded", 7: "conv2^T MMA done", 8: "d_pre written", 9: "depthwise^T done", 10: "conv0^T MMA done",
          11: "dx written, store issued", 12: "B1 MMA done", 13: "dv*u written", 14: "gate-gradient sums done", 15: "teardown done",
          16: "dwd reduced"}
for fold in ("", "both"):
    eng.fold_se = fold
    lvl = int(os.environ.get("LEVEL", 0))
    chain = [op for op in eng.enc_ops[lvl] if hasattr(op, "blocks")][0]
    for _ in range(3):
        eng.forward_train()
    torch.cuda.synchronize()
    buf = torch.zeros(1 + 3000, dtype=torch.int64, device=dev)
    eng._stream()
    lib.mvae_debug_trace(buf.data_ptr())
    chain.fwd()
    torch.cuda.synchronize()
    lib.mvae_debug_trace(0)
    b = buf.cpu().tolist()
    n = b[0]
    ev = [(b[3 + 3 * i], b[1 + 3 * i], b[2 + 3 * i]) for i in range(n)]
    t0 = ev[0][0]
    print(f"--- level {lvl} encoder chain forward, fold_se={fold}: {n} events (tag: 1 = F1, 3 = F2F1, 2 = F2)")
    prev = t0
    for t, e, tag in ev:
        print(f"   {(t - t0) / 1e3:8.2f} us  (+{(t - prev) / 1e3:6.2f})  tag {tag}  {names[e]}")
        prev = t
    if not eng.training:
        continue
    eng.backward()
    torch.cuda.synchronize()
    os.environ["MVAE_DIAG_SKIP_WGRAD"] = "1"       # the tensor-core wgrad kernels write their own records into the buffer
    buf.zero_()
    eng._stream()
    lib.mvae_debug_trace(buf.data_ptr())
    chain.bwd()
    torch.cuda.synchronize()
    lib.mvae_debug_trace(0)
    b = buf.cpu().tolist()
    n = b[0]
    ev = [(b[3 + 3 * i], b[1 + 3 * i], b[2 + 3 * i]) for i in range(n)]
    t0 = ev[0][0]
    os.environ.pop("MVAE_DIAG_SKIP_WGRAD")
    print(f"--- level {lvl} encoder chain backward, fold_se={fold}: {n} events (tag: 11 = B1, 13 = B2B1, 12 = B2)")
    prev = t0
    for t, e, tag in ev:
        print(f"   {(t - t0) / 1e3:8.2f} us  (+{(t - prev) / 1e3:6.2f})  tag {tag}  {bnames[e]}")
        prev = t
