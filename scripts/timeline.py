"""Kernel timeline of one graph-replayed cfg2 training step (CUPTI through torch.profiler): name, stream, start, duration.
Writes gpurun_out/timeline.csv and prints the busy-time summary.  Diagnostic only: numbers under the profiler are not bench values."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from multiscale_variational_autoencoder_b200 import MultiscaleVAE
import bench

cfgname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, B, _ = bench.CONFIGS[cfgname]
dev = torch.device("cuda", 0)
m = MultiscaleVAE(**cfg, precision="tf32", device=dev)
m.compile(0.01, 1.0, 0.1)
if os.environ.get("SERIAL") == "1":
    m.parallel_levels = False
if os.environ.get("FAKE_EARLY") == "1":
    # diagnostic: the graph structure of the early gradient exchange with a one-element kernel in place of the exchange
    import types
    from multiscale_variational_autoencoder_b200 import _lib
    from multiscale_variational_autoencoder_b200.dist import GradAllReduce
    lib = _lib.load()
    scratch = torch.zeros(16, device=dev)
    peer = types.SimpleNamespace(allreduce=lambda stream=None, **kw: _lib.check(lib.mvae_accumulate(
        scratch.data_ptr(), scratch.data_ptr() + 16, 1, 1.0, stream or torch.cuda.current_stream(dev).cuda_stream)),
        timed_out=lambda: False, CHANNELS=16)
    ar = GradAllReduce.__new__(GradAllReduce)
    ar.ps, ar.peer, ar.world, ar.rank = m._ps, peer, 1, 0
    m._dist, m._dp_ingraph, m._dp_early = ar, False, ar.early_ranges()
OUT = os.environ.get("OUT", "timeline")
eng = m._engine(B, True)
H, W, C = cfg["input_dims"]
eng.x.copy_(torch.rand(B, H, W, C) * 255)
for e in eng.eps:
    e.normal_()
for _ in range(5):
    m.train_step_device(eng)
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        m.train_step_device(eng)
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/trace.json")
ev = json.load(open("gpurun_out/trace.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ks.sort(key=lambda e: e["ts"])
# keep the last step: split at the largest gaps
t_end = [e["ts"] + e["dur"] for e in ks]
n = len(ks) // 3
last = ks[2 * n:]
t0 = last[0]["ts"]
with open(f"gpurun_out/{OUT}.csv", "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in last:
        f.write(f'{e["ts"] - t0:.2f},{e["dur"]:.2f},{e["args"].get("stream")},{e["name"][:90]}\n')
span = max(e["ts"] + e["dur"] for e in last) - t0
busy = sum(e["dur"] for e in last)
print(f"kernels/step {len(last)}  span {span:.1f} us  sum of kernel time {busy:.1f} us  avg concurrency {busy / span:.2f}")
per = collections.defaultdict(float)
for e in last:
    per[e["args"].get("stream")] += e["dur"]
for s, d in sorted(per.items(), key=lambda kv: -kv[1]):
    print(f"  stream {s}: {d:.1f} us busy")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in last:
    k = e["name"].split("(")[0][:50]
    agg[k][0] += 1
    agg[k][1] += e["dur"]
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {d:8.1f} us {c:4d} x {k}")
os.remove("gpurun_out/trace.json")
