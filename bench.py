#!/usr/bin/env python
"""Benchmark of the multiscale-VAE training step (BASELINE.json metric: train images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg2]

One "step" = forward + ELBO + backward + regularisers/clipnorm/Adagrad on one synthetic batch.  Default workload:
BASELINE.json configs[1] (32x32x3, 5 levels, batch 256 per GPU, weak scaling under torchrun).

    value        images/s with the batch already resident in HBM (CUDA-graph replay), CUDA events, max over ranks
    e2e          images/s through MultiscaleVAE.train_on_batch-equivalent calls: pinned-host -> device copy of the batch
                 and of eps every step, device -> host read of the loss every step, inside the timed region
    roofline     the dominant kernel launch of the step (one eager, single-stream pass with CUDA events around every
                 C-ABI call), algorithmic bytes / FLOPs per launch as defined in DESIGN.md
    cpu_baseline the CPU restatement of the reference step (oracle/, PyTorch-CPU fp32, all host threads) on a bounded
                 sample of the same workload; --impl reference prints that arm alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (model kwargs, per-GPU batch, description)
    "cfg1": (dict(input_dims=(32, 32, 3), z_dims=[128, 64, 32], sample_std=0.5,
                  encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (2, 2), (1, 1)]}),
             32, "CIFAR-10-shaped 32x32x3, 3 levels, batch 32 (main.py:81-91)"),
    "cfg2": (dict(input_dims=(32, 32, 3), z_dims=[128, 64, 32, 16, 8], sample_std=0.5,
                  encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (1, 1), (1, 1)]}),
             256, "CIFAR-10-shaped 32x32x3, 5 levels (full log2 depth), batch 256 per GPU"),
    "cfg3": (dict(input_dims=(64, 64, 3), z_dims=[128, 64, 32, 16, 8, 8], sample_std=0.5,
                  encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (1, 1), (1, 1)]}),
             64, "64x64x3, 6 levels, global batch 512 at 8 GPUs"),
    "cfg4": (dict(input_dims=(256, 256, 3), z_dims=[32, 32, 32, 32, 16, 16, 8, 8], sample_std=0.5,
                  encoder={"filters": [64, 128, 128], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (1, 1), (1, 1)]}),
             64, "256x256x3, 8 levels, batch 64 per GPU, wide filters [64,128,128]"),
}
LR, RF, KF = 0.01, 1.0, 0.1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1400.0, src="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                  "-lms", "20"], stdout=subprocess.PIPE, text=True)
        except OSError:
            return
        while not self.stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        time.sleep(0.15)
        self.stop.set()
        self.t.join(2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------------
class ProfilingLib:
    """Wraps the ctypes library: CUDA events around every C-ABI call + algorithmic bytes / FLOPs of that call."""

    def __init__(self, lib, torch):
        self._lib, self._torch, self.records = lib, torch, []

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("mvae_") or name.endswith("_bytes") or name.startswith("mvae_set_"):
            return fn                      # host-side queries / policy setters launch nothing
        torch = self._torch

        def wrapped(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, args, e0, e1))
            return rc

        return wrapped


def account(name, args):
    """(shape key, algorithmic bytes, FLOPs) of one C-ABI call; definitions in DESIGN.md section 5."""
    if name in ("mvae_conv2d_fwd", "mvae_conv2d_dgrad", "mvae_conv2d_wgrad"):
        d = args[0]._obj
        Ho, Wo = -(-d.H // d.sh), -(-d.W // d.sw)
        cin_t = d.Cin + d.coord_mode
        xin, yout, wsz = d.B * d.H * d.W * d.Cin, d.B * Ho * Wo * d.Cout, d.kh * d.kw * cin_t * d.Cout
        flops = 2.0 * d.B * Ho * Wo * d.kh * d.kw * cin_t * d.Cout
        extra = 0
        if name == "mvae_conv2d_fwd" and args[5]:
            extra += yout                      # residual read
        if name == "mvae_conv2d_dgrad":
            extra += xin * (bool(args[4]) + bool(args[5]))   # residual, act_out reads
        key = f"B{d.B} {d.H}x{d.W}x{d.Cin}->{d.Cout} k{d.kh} s{d.sh}"
        return key, 4.0 * (xin + yout + wsz + extra), flops
    if name in ("mvae_dense_fwd", "mvae_dense_dgrad"):
        M, K, N = args[0:3]
        extra = M * K if (name == "mvae_dense_dgrad" and args[5]) else 0       # activation-output read
        return f"M{M} K{K} N{N}", 4.0 * (M * K + K * N + M * N + extra), 2.0 * M * K * N
    if name == "mvae_conv2d_wgrad_batched":
        # the deferred weight gradients of one level: n problems of one layer shape in one launch
        n, ds = args[0], args[1]
        by = fl = 0.0
        shapes = {}
        for i in range(n):
            d = ds[i]
            Ho, Wo = -(-d.H // d.sh), -(-d.W // d.sw)
            xin, yout, wsz = d.B * d.H * d.W * d.Cin, d.B * Ho * Wo * d.Cout, d.kh * d.kw * d.Cin * d.Cout
            by += 4.0 * (xin + yout + wsz)
            fl += 2.0 * d.B * Ho * Wo * d.kh * d.kw * d.Cin * d.Cout
            shapes[f"{d.H}x{d.W}"] = shapes.get(f"{d.H}x{d.W}", 0) + 1
        d = ds[0]
        key = f"B{d.B} " + "+".join(f"{c}@{k}" for k, c in sorted(shapes.items())) + f" x{d.Cin}->{d.Cout} k{d.kh} s{d.sh}"
        return key, by, fl
    if name == "mvae_dwconv3x3_fwd":
        B, H, W, Cc = args[5:9]
        return f"B{B} {H}x{W}x{Cc}", 4.0 * 2 * B * H * W * Cc, 2.0 * 9 * B * H * W * Cc
    if name == "mvae_dwconv3x3_bwd":
        B, H, W, Cc = args[9:13]
        return f"B{B} {H}x{W}x{Cc}", 4.0 * 4 * B * H * W * Cc, 2.0 * 27 * B * H * W * Cc
    if name == "mvae_se_dgate_reduce":
        B, HW, Cc = args[3:6]
        return f"B{B} {HW}x{Cc}", 4.0 * 2 * B * HW * Cc, 2.0 * B * HW * Cc
    if name == "mvae_bn_stats":
        M, Cc = args[2:4]
        return f"M{M}x{Cc}", 4.0 * M * Cc, 3.0 * M * Cc
    if name == "mvae_colsum":
        M, Cc = args[2:4]
        return f"M{M}x{Cc}", 4.0 * M * Cc, 1.0 * M * Cc
    if name == "mvae_bn_convout_fwd":
        M, Cf, Co = args[10:13]
        return f"M{M} {Cf}->{Co}", 4.0 * M * (Cf + Co), 2.0 * M * Cf * Co
    if name == "mvae_bn_convout_bwd":
        M, Cf, Co = args[12:15]
        return f"M{M} {Cf}->{Co}", 4.0 * M * (2 * Cf + Co), 6.0 * M * Cf * Co
    if name == "mvae_pyramid_split":
        B, H, W, Cc, L = args[3:8]
        n0 = B * H * W * Cc
        return f"B{B} {H}x{W}x{Cc} L{L}", 4.0 * (n0 + sum(n0 >> (2 * i) for i in range(L))), 18.0 * n0
    if name in ("mvae_pyramid_merge_fwd", "mvae_pyramid_merge_bwd"):
        B, H, W, Cc, L = (args[3:8] if name.endswith("fwd") else args[2:7])
        n0 = B * H * W * Cc
        alias = name.endswith("bwd") and args[0] == args[1][0]      # dys[0] is dr0: nothing to copy
        return (f"B{B} {H}x{W}x{Cc} L{L}", 4.0 * ((0 if alias else n0) + sum(n0 >> (2 * i) for i in range(L))), 8.0 * n0)
    if name == "mvae_recon_loss_fwd":
        B, H, W, Cc = args[4:8]
        return f"B{B} {H}x{W}x{Cc}", 4.0 * (3 if args[2] else 2) * B * H * W * Cc, 6.0 * B * H * W * Cc
    if name == "mvae_recon_loss_bwd":
        B, H, W, Cc = args[4:8]
        return f"B{B} {H}x{W}x{Cc}", 4.0 * 3 * B * H * W * Cc, 8.0 * B * H * W * Cc
    if name in ("mvae_mbv3_fused_fwd", "mvae_mbv3_fused_bwd"):
        # fused mobilenetV3 tile kernels: every tensor of a phase crosses HBM once
        a = args[0]._obj
        n = a.B * a.H * a.W * a.C
        conv = 2.0 * a.B * a.H * a.W * a.C * a.C
        if name.endswith("fwd"):
            f2, f1 = bool(a.w2), bool(a.w0)
            tensors = (3 if f2 else 0) + ((1 + bool(a.a) + (0 if f2 else 1)) if f1 else 0)
            flops = conv * (f2 + f1) + (18.0 * n if f1 else 0.0)
            tag = ("F2" if f2 else "") + ("F1" if f1 else "")
        else:
            b2, b1 = bool(a.w0), bool(a.w2_prev)
            tensors = (5 if b2 else 0) + ((1 + (0 if b2 else 1)) if b1 else 0)
            flops = conv * (2 * b2 + b1) + (36.0 * n if b2 else 0.0)
            tag = ("B2" if b2 else "") + ("B1" if b1 else "")
        return f"B{a.B} {a.H}x{a.W}x{a.C} {tag}", 4.0 * tensors * n, flops
    if name == "mvae_se_gate_fwd":
        B, Cc = args[11:13]
        return f"B{B} C{Cc}", 4.0 * (2 * B * Cc + 2 * Cc * Cc), 4.0 * B * Cc * Cc
    if name == "mvae_se_gate_bwd":
        B, Cc = args[13:15]
        return f"B{B} C{Cc}", 4.0 * (2 * B * Cc + 4 * Cc * Cc), 12.0 * B * Cc * Cc
    return "", 0.0, 0.0


def profile_step(model, eng, torch):
    """One eager, single-stream step with events around every C-ABI call (shares; ncu launch lists are the cross-check),
    then the steady-state device time of the heaviest calls: each is captured 20x back to back into a CUDA graph with its
    real arguments and replayed between two events, so neither host launch gaps nor event overhead are in the number.
    Returns per-(call, shape) aggregates with `us` = per-launch device time (graph-replay where measured)."""
    from multiscale_variational_autoencoder_b200 import _lib
    real = eng.lib
    prof = ProfilingLib(real, torch)
    eng.lib = prof
    try:
        torch.cuda._sleep(int(40e6))      # ~20 ms: the step is enqueued while the GPU spins -> few host gaps
        eng.defer_serial = True           # weight gradients as the batched launches the captured graph issues
        eng.forward_backward(parallel=False)
        eng.optimizer_step(model._lr_dev, model._clip_norm, 1.0 / model._world)
        torch.cuda.synchronize()
    finally:
        eng.lib = real
        eng.defer_serial = False
    agg = {}
    for name, args, e0, e1 in prof.records:
        ms = e0.elapsed_time(e1)
        key, by, fl = account(name, args)
        a = agg.setdefault((name, key), dict(calls=0, ms=0.0, bytes=by, flops=fl, args=args))
        a["calls"] += 1
        a["ms"] += ms
    # steady-state re-measurement of the top groups
    top = sorted((k for k, v in agg.items() if v["bytes"] > 0), key=lambda k: -agg[k]["ms"])[:16]
    st = torch.cuda.Stream()
    for k in top:
        v = agg[k]
        fn, args = getattr(real, k[0]), list(v["args"])
        with torch.cuda.stream(st):
            args[-1] = st.cuda_stream
            for _ in range(2):
                _lib.check(fn(*args), k[0])
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                args[-1] = torch.cuda.current_stream().cuda_stream
                for _ in range(20):
                    _lib.check(fn(*args), k[0])
            g.replay()
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            st.synchronize()
        v["us_replay"] = e0.elapsed_time(e1) / 20 * 1e3
    for v in agg.values():
        v.pop("args")
        v["us"] = v.get("us_replay", v["ms"] / v["calls"] * 1e3)
    return agg, len(prof.records)


# ----------------------------------------------------------------------------------------------------------------------
def hbm_microbench(torch, device, B=128, H=512, W=512, C=3, L=9, zdims=(128, 64, 32, 16, 8, 8, 8, 8, 8), iters=10, warm=3):
    """BASELINE.json configs[4]: pyramid split / merge fwd / merge bwd / reconstruction loss fwd+bwd / reparam+KL on
    512x512x3, batch 128, 9 levels.  Every tensor set is > 126 MB (L2), so successive launches see cold DRAM.
    Returns {kernel: {us, algorithmic_bytes, GB/s, frac}} with the algorithmic byte counts of DESIGN.md section 5."""
    import ctypes as Ct
    from multiscale_variational_autoencoder_b200 import _lib
    from multiscale_variational_autoencoder_b200.engine import gaussian_kernel
    import numpy as np
    lib = _lib.load()
    pk = peaks()
    s = torch.cuda.current_stream(device).cuda_stream
    f32 = dict(dtype=torch.float32, device=device)
    n0 = B * H * W * C
    lv = [B * (H >> i) * (W >> i) * C for i in range(L)]
    x = torch.rand(B, H, W, C, **f32) * 255
    bands = [torch.empty(B, H >> i, W >> i, C, **f32) for i in range(L)]
    ys = [torch.randn(B, H >> i, W >> i, C, **f32) * 0.1 for i in range(L)]
    dys = [torch.empty(B, H >> i, W >> i, C, **f32) for i in range(L)]
    r0 = torch.empty(B, H, W, C, **f32)
    sums = torch.zeros(B * (1 + 2 * C), **f32)
    ws_s = torch.empty(lib.mvae_pyramid_split_workspace_bytes(B, H, W, C, L) // 4 + 1, **f32)
    ws_m = torch.empty(lib.mvae_pyramid_merge_workspace_bytes(B, H, W, C, L) // 4 + 1, **f32)
    taps = (Ct.c_float * 9)(*[float(v) for v in gaussian_kernel((3, 3), (2, 2)).astype(np.float32).ravel()])
    P = lambda ts: (Ct.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    bp, yp, dyp = P(bands), P(ys), P(dys)
    zt = sum(zdims)
    mulv = torch.randn(B, 2 * zt, **f32) * 0.1
    eps = torch.randn(B, zt, **f32)
    z = torch.empty(B, zt, **f32)
    dz = torch.randn(B, zt, **f32)
    dmulv = torch.empty(B, 2 * zt, **f32)
    kl = torch.empty(B, **f32)
    pyr = 4.0 * (n0 + sum(lv))
    cases = [
        ("pyramid_split", pyr, lambda: lib.mvae_pyramid_split(x.data_ptr(), bp, ws_s.data_ptr(), B, H, W, C, L, 0.0, 255.0,
                                                              taps, 3, 3, 0, s)),
        ("pyramid_merge_fwd", pyr, lambda: lib.mvae_pyramid_merge_fwd(yp, r0.data_ptr(), ws_m.data_ptr(), B, H, W, C, L, s)),
        ("recon_loss_fwd", 8.0 * n0, lambda: lib.mvae_recon_loss_fwd(r0.data_ptr(), x.data_ptr(), 0, sums.data_ptr(), B, H, W,
                                                                     C, 0.0, 255.0, s)),
        ("recon_loss_bwd", 12.0 * n0, lambda: lib.mvae_recon_loss_bwd(r0.data_ptr(), x.data_ptr(), sums.data_ptr(),
                                                                      dys[0].data_ptr(), B, H, W, C, 0.0, 255.0, 1.0 / B, s)),
        # dys[0] aliases dr0 (no copy): read dr0 once, write the L-1 coarser gradients
        ("pyramid_merge_bwd", 4.0 * sum(lv), lambda: lib.mvae_pyramid_merge_bwd(dys[0].data_ptr(), dyp, B, H, W, C, L, s)),
        ("reparam_kl_fwd", 4.0 * (4 * B * zt + B), lambda: lib.mvae_reparam_kl_fwd(mulv.data_ptr(), eps.data_ptr(), z.data_ptr(),
                                                                                   kl.data_ptr(), B, zt, 1.0, 0.5, s)),
        ("reparam_kl_bwd", 4.0 * 6 * B * zt, lambda: lib.mvae_reparam_kl_bwd(mulv.data_ptr(), eps.data_ptr(), dz.data_ptr(),
                                                                             dmulv.data_ptr(), B, zt, 1.0, 0.5, 0.1 / B, s)),
    ]
    out = {}
    if warm:
        # the CPU baseline ran just before (GPU idle for seconds): bring the clocks back up before anything is timed
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            for name, _, fn in cases[:5]:
                _lib.check(fn(), name)
            torch.cuda.synchronize()
    for name, nbytes, fn in cases:
        for _ in range(warm):
            _lib.check(fn(), name)
        evs = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)          # ~0.2 ms GPU spin: the launches below are queued before e0 fires (no host gaps)
            e0.record()
            _lib.check(fn(), name)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in evs)
        us = t[len(t) // 2] * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        out[name] = dict(us=round(us, 2), algorithmic_bytes=nbytes, gbs=round(gbs, 1), frac=round(gbs / pk["hbm"], 4))
    del x, bands, ys, dys, r0
    torch.cuda.empty_cache()
    return dict(workload=f"cfg5: {H}x{W}x{C} batch {B}, {L} levels (BASELINE.json configs[4])", peak_gbs=pk["hbm"],
                peak_source=pk["src"], timing=f"median of {iters} calls, CUDA events, each call queued behind a GPU spin (no host gaps), operands > L2",
                kernels=out)

# ----------------------------------------------------------------------------------------------------------------------
def cpu_step_rate(cfg, B, steps, warmup):
    """CPU restatement of the reference training step (the oracle; TensorFlow/Keras cannot run here)."""
    import torch
    from oracle.mvae_oracle import OracleMVAE
    torch.set_num_threads(os.cpu_count() or 1)
    m = OracleMVAE(dtype=torch.float32, **cfg)
    m.compile(LR, RF, KF)
    g = torch.Generator().manual_seed(1234)
    H, W, C = cfg["input_dims"]
    x = torch.rand(B, H, W, C, generator=g) * 255
    eps = [torch.randn(B, z, generator=g) for z in cfg["z_dims"]]
    for _ in range(warmup):
        m.train_step(x, eps)
    t0 = time.perf_counter()
    for _ in range(steps):
        m.train_step(x, eps)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def workload_config(name, desc, cfg, B, world, precision):
    """The `config` object of the JSON line; both arms print the same keys."""
    return dict(workload=f"{name}: {desc}", per_gpu_batch=B, global_batch=B * world, levels=len(cfg["z_dims"]),
                parallelism=f"dp{world}", optimizer="adagrad+clipnorm+l1/l2", eps="supplied per step", l2="", precision=precision)


def time_fit_like(torch, model, eng, steps, warmup):
    """The step as `train()` runs it (multiscale_vae.py:139-147 + :550-557): GaussianNoise + SpatialDropout2D of the batch
    and eps / noise / dropout masks drawn by the device RNG every step, in front of the graph-replayed step."""
    def one():
        model._load_eps(eng, None)
        model._corrupt(eng)
        model.train_step_device(eng)
    ce = model._engine(eng.B, True, True)
    ce.x.copy_(eng.x)
    eng = ce
    for _ in range(max(warmup, 3)):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(ms_per_step=ms, images_per_s=eng.B / (ms / 1e3),
                note="training-phase input corruption (GaussianNoise + SpatialDropout2D) and device-drawn eps every step, as "
                     "train() / Keras fit run it; the headline supplies eps and leaves the corruption out")


def time_config(torch, dist, name, precision, device, world, steps, warmup):
    """Graph-replayed training step of another BASELINE config (per-GPU batch of CONFIGS): ms/step (max over ranks),
    images/s, tensor-core FLOP/s of the convolutions / Dense layers (3x forward FLOPs minus conv_base dgrad)."""
    from multiscale_variational_autoencoder_b200 import MultiscaleVAE
    cfg, B, desc = CONFIGS[name]
    model = MultiscaleVAE(**cfg, precision=precision, device=device)
    model.compile(LR, RF, KF)
    if world > 1:
        model.enable_data_parallel()
    eng = model._engine(B, True)
    H, W, C = cfg["input_dims"]
    g = torch.Generator().manual_seed(99)
    eng.x.copy_(torch.rand(B, H, W, C, generator=g) * 255)
    for e in eng.eps:
        e.copy_(torch.randn(e.shape, generator=g))
    if world > 1:
        # the gradient exchange is a kernel that waits for its peers on the device: every rank has built its engine (tens
        # of GB of buffers at cfg4) before the first step runs anywhere
        torch.cuda.synchronize()
        dist.barrier()
    for _ in range(max(warmup, 3)):
        model.train_step_device(eng)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.train_step_device(eng)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    flops = step_tensor_flops(eng)
    pk = peaks()
    peak_tf = pk["tf"] / 2 if precision == "tf32" else None
    out = dict(workload=f"{name}: {desc}", per_gpu_batch=B, global_batch=B * world, ms_per_step=ms,
               images_per_s=B * world / (ms / 1e3), steps=steps, tensor_gflop_per_step=flops / 1e9,
               tensor_tflops=flops / (ms * 1e-3) / 1e12)
    if peak_tf:
        out["frac_tensor_tf32_derived"] = out["tensor_tflops"] / peak_tf
    del model, eng
    torch.cuda.empty_cache()
    return out


def step_tensor_flops(eng):
    """FLOPs of the dense contractions of one training step of this engine (fwd + dgrad + wgrad of every Conv2D /
    Conv2DTranspose / Dense, no dgrad for conv_base): the tensor-roofline numerator of SURVEY 8(d)."""
    total = 0.0
    for ops in eng.enc_ops + eng.dec_ops:
        for op in [b for o in ops for b in getattr(o, "blocks", [o])]:
            for dn, need_dx in (("desc", getattr(op, "need_dx", True)), ("d0", True), ("d2", True)):
                d = getattr(op, dn, None)
                if d is None:
                    continue
                Ho, Wo = -(-d.H // d.sh), -(-d.W // d.sw)
                f = 2.0 * d.B * Ho * Wo * d.kh * d.kw * (d.Cin + d.coord_mode) * d.Cout
                total += f * (3 if need_dx else 2)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--extra-configs", default="cfg1,cfg3,cfg4", help="other BASELINE configs timed next to the headline")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--precision", default=os.environ.get("MVAE_PRECISION", "tf32"), choices=["fp32", "tf32"],
                    help="tf32: tcgen05 tensor cores (north-star tolerance 1e-3); fp32: CUDA-core kernels (1e-5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches on one stream (for ncu launch lists)")
    ap.add_argument("--micro-only", action="store_true", help="run only the cfg5 pyramid/ELBO HBM microbenchmark")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--micro-iters", type=int, default=10, help="1 = single cold launch per kernel (for ncu captures)")
    ap.add_argument("--profile-json", default="", help="write the per-kernel table of the profiling pass here")
    a = ap.parse_args()
    cfg, B, desc = CONFIGS[a.config]
    if a.batch:
        B = a.batch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = workload_config(a.config, desc, cfg, B, world, a.precision)

    # ------------------------------------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU path (oracle restatement: TensorFlow cannot run here) on ONE process with all host threads;
        # --steps / --warmup are honoured up to a time bound, each step a full per-GPU batch of the workload when a step
        # takes well under a second, else a bounded sample of it
        sample_b = B if a.config in ("cfg1", "cfg2") else min(B, 8 if a.config == "cfg3" else 2)
        probe_ips, probe_ms, cores = cpu_step_rate(cfg, sample_b, 1, 1)
        budget_s = 150.0
        steps = max(1, min(a.steps, int(budget_s / (probe_ms / 1e3))))
        warm = max(1, min(a.warmup, max(1, int(10.0 / (probe_ms / 1e3)))))
        ips, ms, cores = cpu_step_rate(cfg, sample_b, steps, warm)
        workload.update(global_batch=sample_b, per_gpu_batch=sample_b, parallelism="cpu (1 process, all host threads)",
                        l2="n/a (CPU arm)", precision="f32")
        note = "" if (steps, warm) == (a.steps, a.warmup) else \
            f"asked for --steps {a.steps} --warmup {a.warmup}; ran {steps}/{warm} to stay inside {budget_s:.0f} s of CPU time"
        print(json.dumps(dict(
            impl="reference", metric="train images/sec", value=ips, unit="images/s", n_gpus=a.gpus, steps=steps,
            warmup=warm, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
            data="synthetic", config=workload, note=note,
            cpu_baseline=dict(value=ips, unit="images/s", cores=cores, kind="port",
                              sample=f"{steps} steps of batch {sample_b} after {warm} warm-ups (oracle/mvae_oracle.py: CPU "
                                     "restatement of the reference step; TensorFlow 2.3.1/Keras 2.4.3 are not installable here)"),
            e2e=dict(value=ips, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return

    # ----------------------------------------------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from multiscale_variational_autoencoder_b200 import MultiscaleVAE
    if a.micro_only:
        print(json.dumps(hbm_microbench(torch, device, iters=a.micro_iters, warm=0 if a.micro_iters == 1 else 3)))
        return
    model = MultiscaleVAE(**cfg, precision=a.precision, device=device)
    model.compile(LR, RF, KF)
    if a.no_graph:
        model.use_cuda_graph = model.parallel_levels = False
    if os.environ.get("MVAE_SERIAL_LEVELS") == "1":      # diagnostic: CUDA graph, but all levels on one stream
        model.parallel_levels = False
    if world > 1:
        model.enable_data_parallel()
    eng = model._engine(B, True)
    H, W, C = cfg["input_dims"]
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = (torch.rand(B, H, W, C, generator=g) * 255).pin_memory()
    eps_host = [torch.randn(B, z, generator=g).pin_memory() for z in cfg["z_dims"]]
    loss_host = torch.zeros(5).pin_memory()
    eng.x.copy_(x_host)
    for e, h in zip(eng.eps, eps_host):
        e.copy_(h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # --- device-resident throughput -------------------------------------------------------------------------------
    barrier()          # every rank has built its engine before the first step (the exchange kernel waits for its peers)
    for _ in range(max(a.warmup, 3)):
        model.train_step_device(eng)
    barrier()
    with ClockSampler(local) as clk:
        # the same load runs (untimed) before and after the timed K steps, so that the 20 ms clock samples bracket the
        # timed region even when K steps last only a few tens of milliseconds
        n_pad = {"cfg1": 200, "cfg2": 100, "cfg3": 80, "cfg4": 3}[a.config]      # a fixed count: every rank must take the same steps
        for _ in range(n_pad):
            model.train_step_device(eng)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(a.steps):
            model.train_step_device(eng)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        for _ in range(n_pad):
            model.train_step_device(eng)
        torch.cuda.synchronize()
    clocks = clk.summary()
    value = B * world * a.steps / (ms / 1e3)
    kernels_per_step = getattr(eng, "kernels_per_step", 0)

    # --- end to end: host buffers in, loss out, every step ------------------------------------------------------------
    # the library's input pipeline (MultiscaleVAE.stage_batch / train_step_staged): the H2D copy of step i+1 runs on a copy
    # stream while step i computes; every step's inputs still cross PCIe inside the timed region, and the loss is read
    # back (and waited for) every step
    model.stage_batch(eng, x_host, eps_host)

    # The loss of every step is copied to the host and read there; the read of step i happens after step i+1 has been
    # enqueued (one step of lag, as a training loop that logs its loss does), so the host's launch work is off the GPU's
    # critical path.
    loss_slots = [torch.zeros(5).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    state = dict(i=0, pending=None)

    def read_pending():
        j = state["pending"]
        if j is None:
            return None
        loss_ev[j].synchronize()
        state["pending"] = None
        return float(loss_slots[j][0] + loss_slots[j][4])

    def e2e_step():
        j = state["i"] & 1
        state["i"] += 1
        model.stage_batch(eng, x_host, eps_host)          # next step's inputs
        model.train_step_staged(eng)
        loss_slots[j].copy_(eng.loss5, non_blocking=True)      # the five loss scalars sit side by side: one D2H copy
        loss_ev[j].record()
        prev = read_pending()                             # loss of the previous step (this step is already queued)
        state["pending"] = j
        return prev

    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        last_loss = e2e_step()
    last_loss = read_pending()                            # the last step's loss is read inside the timed region too
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    h2d = x_host.numel() * 4 + sum(h.numel() * 4 for h in eps_host)
    e2e = dict(value=B * world * a.steps / (ms_e2e / 1e3), unit="images/s", h2d_bytes_per_step=h2d,
               d2h_bytes_per_step=20, ms_per_step=ms_e2e / a.steps, api="MultiscaleVAE.stage_batch + train_step_staged (pinned H2D prefetch), loss D2H every step, read one step late")

    # --- per-kernel pass: dominant launch and its roofline ----------------------------------------------------------------
    pk = peaks()
    profile_step(model, eng, torch)                      # warm (eager path)
    agg, launches = profile_step(model, eng, torch)
    tot = sum(v["ms"] for v in agg.values())
    # dominant KERNEL = the C-ABI call with the largest total device time over all its launches of the step (all shapes:
    # this is how the ncu launch list under profiles/ groups them); its roofline uses the launch-weighted averages
    # (sum of algorithmic bytes / sum of device time), and the heaviest single shape is reported next to it
    by_call = {}
    for (name, key), v in agg.items():
        if v["bytes"] <= 0:
            continue
        c = by_call.setdefault(name, dict(calls=0, us=0.0, bytes=0.0, flops=0.0, shapes=[]))
        c["calls"] += v["calls"]
        c["us"] += v["calls"] * v["us"]
        c["bytes"] += v["calls"] * v["bytes"]
        c["flops"] += v["calls"] * v["flops"]
        c["shapes"].append((v["calls"] * v["us"], key, v))
    dname, dc = max(by_call.items(), key=lambda kv: kv[1]["us"])
    _, dkey, hv = max(dc["shapes"], key=lambda t: t[0])
    dv = dict(calls=dc["calls"], us=dc["us"] / dc["calls"], bytes=dc["bytes"] / dc["calls"], flops=dc["flops"] / dc["calls"])
    per_ms = dv["us"] / 1e3
    ai = dv["flops"] / max(dv["bytes"], 1.0)
    ridge = pk["tf"] * 1e12 / (pk["hbm"] * 1e9)
    fp32_note = ""
    if ai > ridge:
        bound, ach, peak, unit = "tensor", dv["flops"] / (per_ms * 1e-3) / 1e12, pk["tf"], "TFLOP/s"
        if a.precision == "tf32":
            peak, fp32_note = pk["tf"] / 2, " (TF32 peak derived as half of the measured bf16 figure)"
    else:
        bound, ach, peak, unit = "hbm", dv["bytes"] / (per_ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{dname} {dkey}")
    step_us = sum(v["calls"] * v["us"] for v in agg.values())
    hv_ach = hv["bytes"] / (hv["us"] * 1e-6) / 1e9
    roofline = dict(bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, traffic=traffic,
                    kernel=f"{dname} [all {len(dc['shapes'])} shapes of the step]", calls_per_step=dv["calls"],
                    ms_per_launch=per_ms, share_of_step=dv["calls"] * dv["us"] / step_us, arithmetic_intensity=ai,
                    peak_source=pk["src"] + fp32_note, algorithmic_bytes_per_launch=dv["bytes"],
                    algorithmic_flops_per_launch=dv["flops"],
                    heaviest_shape=dict(shape=dkey, calls=hv["calls"], us_per_launch=hv["us"],
                                        algorithmic_bytes=hv["bytes"], gbs=hv_ach, frac_hbm=hv_ach / pk["hbm"],
                                        traffic_note="`traffic` is the ncu DRAM byte count of this shape's launch"),
                    timing="20 back-to-back launches with the step's real arguments, CUDA-graph replay between two events "
                           "(L2-warm: operands were just produced, as in the step)",
                    note="single weight-gradient launches are limited to a 32-SM share by design (they run beside the dgrad "
                         "chain, MVAE_WGRAD_SMS); the deferred ones go out as batched launches on all SMs "
                         "(mvae_conv2d_wgrad_batched, eight 1x1 problems: 3.7 TB/s)" if "wgrad" in dname else "")
    if a.profile_json and rank == 0:
        rows = sorted(({"call": k[0], "shape": k[1], **v, "share": v["calls"] * v["us"] / step_us} for k, v in agg.items()),
                      key=lambda r: -r["calls"] * r["us"])
        with open(a.profile_json, "w") as f:
            json.dump(dict(config=a.config, batch=B, precision=a.precision, eager_step_ms=tot, rows=rows), f, indent=1)

    # --- CPU baseline (rank 0, N=1 only) ----------------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        ips, cms, cores = cpu_step_rate(cfg, B, 8, 2)
        cpu = dict(value=ips, unit="images/s", cores=cores, kind="port", ms_per_step=cms,
                   sample=f"8 steps of batch {B} after 2 warm-ups (oracle/mvae_oracle.py, PyTorch-CPU fp32)")

    micro = None
    if rank == 0 and world == 1 and not a.no_micro:
        micro = hbm_microbench(torch, device)

    # the parity-tight precision next to the benchmarked one, and the other BASELINE configs (every rank takes part: the
    # data-parallel exchange is inside the step)
    act_mb = sum(t.numel() * 4 for ops in eng.enc_ops + eng.dec_ops for op in ops for t in [getattr(op, "y").data]) / 1e6
    other = {}
    exchange = None
    if world > 1:
        peer = model._dist.peer
        if peer is not None and peer.timed_out():
            raise RuntimeError("the peer-memory gradient exchange timed out on a barrier (a rank stopped?)")
        exchange = ("peer-memory kernel over NVLink (mvae_comm_allreduce, captured in the step graph; the big Dense "
                    "gradients are exchanged behind their weight-gradient launches, the rest at the end of the backward pass)" if peer is not None else "NCCL all-reduce between the two graphs")
    if not a.no_extra:
        if a.precision == "tf32" and world == 1:
            other["fit_like_same_config"] = time_fit_like(torch, model, eng, 100, 5)
        del model, eng
        torch.cuda.empty_cache()
        if a.precision == "tf32":
            r = time_config(torch, dist if world > 1 else None, a.config, "fp32", device, world, 50, 5)
            other["fp32_same_config"] = dict(ms_per_step=r["ms_per_step"], images_per_s=r["images_per_s"],
                                             note="precision='fp32': CUDA-core kernels, 1e-5 parity")
        for name in [c for c in a.extra_configs.split(",") if c and c != a.config]:
            st = {"cfg1": 200, "cfg2": 100, "cfg3": 60, "cfg4": 6}[name]
            other[name] = time_config(torch, dist if world > 1 else None, name, a.precision, device, world, st, 5)

    if rank == 0:
        workload["l2"] = f"no flush: one step streams > {act_mb:.0f} MB of activations (+ gradients), L2 is 126 MB"
        workload["precision"] = a.precision
        print(json.dumps(dict(
            metric="train images/sec", value=value, unit="images/s", n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
            ms_per_step=ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32" if a.precision == "fp32" else "tf32", data="synthetic", config=workload, clocks=clocks, e2e=e2e,
            gpu_launches=(kernels_per_step or launches) * a.steps, launches_per_step=kernels_per_step or launches,
            launches_note="kernel nodes of one replay of the captured step (library launch counter around the capture); "
                          f"the eager per-call pass made {launches} C-ABI calls",
            roofline=roofline, cpu_baseline=cpu,
            pyramid_elbo_hbm=micro, configs=other, last_loss=last_loss, grad_exchange=exchange)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
