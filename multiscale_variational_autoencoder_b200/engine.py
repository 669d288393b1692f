"""Executor of the multiscale-VAE graph on one B200.

`ParamStore` keeps every variable of the model in ONE flat fp32 buffer (gradients and Adagrad accumulators mirror
its layout), named and laid out like the Keras variables of the reference.  `Engine` is the
per-batch-size plan: it owns the activation buffers and an ordered list of ops whose `fwd`/`bwd` enqueue the
hand-written kernels of libmvae_b200.so through the C-ABI (include/mvae_b200.h).  PyTorch supplies device memory,
streams and CUDA-graph capture only -- there is no torch arithmetic on the step path.

Reference graph: mvae/multiscale_vae.py:129-160 (pyramid), :319-385 (encoder), :389-433 (decoder), :204-224 (merge),
:453-499 (loss, optimiser); mvae/layer_blocks.py:418-462, 556-648, 893-974 (blocks).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import (ACT_ELU, ACT_NONE, ACT_RELU, DIFF_LAPLACIAN, DIFF_NO_UPSAMPLE, PREC_FP32, PREC_TF32, REG_L1, REG_L2,
                   REG_NONE, ConvDesc, Mbv3BwdArgs, Mbv3FwdArgs, check)

SE_BN_EPS, SE_BN_MOM = 1e-3, 0.99          # Keras BatchNormalization defaults (layer_blocks.py:447-449)
DEC_BN_EPS, DEC_BN_MOM = 1e-4, 0.999       # multiscale_vae.py:420-421
CHUNK = 2048
ALIGN = 64                                  # floats; keeps every variable 256-byte aligned


def gaussian_kernel(size, nsig):
    """Host constant of layer_blocks.py:980-1002 (fp64), same expression (distance first, then its square) so that the taps
    equal the reference function's output bit for bit (tests/golden/gaussian_kernel.npz, tests/test_host.py)."""
    assert len(nsig) == 2 and len(size) == 2
    axes = [np.linspace(-abs(nsig[i]), abs(nsig[i]), size[i], endpoint=True) for i in range(2)]
    gx, gy = np.meshgrid(axes[0], axes[1])
    d = np.sqrt(gx * gx + gy * gy)
    g = np.exp(-(d ** 2) / 2.0)
    return g / g.sum()


def same_out(size, s):
    return -(-size // s)


# =============================================================================================================
# Parameters
# =============================================================================================================
class ParamStore:
    def __init__(self, device, seed=7):
        self.device = device
        self.entries = OrderedDict()      # name -> dict(offset, shape, reg, trainable, init)
        self.aliases = OrderedDict()      # keras name -> (storage name, column slice)
        self.segs = []                    # (offset, count, width, ld, reg, keras name)
        self.size = 0
        self.gen = torch.Generator().manual_seed(seed)
        self.flat = self.grads = self.acc = None

    def add(self, name, shape, reg=REG_NONE, fans=None, value=0.0, trainable=True, seg=True):
        n = int(np.prod(shape))
        self.entries[name] = dict(offset=self.size, shape=tuple(shape), reg=reg, trainable=trainable, fans=fans,
                                  value=value)
        if trainable and seg:
            self.segs.append((self.size, n, n, n, reg, name))
        self.size += -(-n // ALIGN) * ALIGN
        return name

    def add_fused_pair(self, name, rows, z, reg, names, fans):
        """(rows, 2z) storage exposing two Keras variables (rows, z) as column halves (mu | log_var)."""
        n = rows * 2 * z
        self.entries[name] = dict(offset=self.size, shape=(rows, 2 * z), reg=reg, trainable=True, fans=fans, value=0.0,
                                  pair=z)
        for h, kn in enumerate(names):
            self.segs.append((self.size + h * z, rows * z, z, 2 * z, reg, kn))
            self.aliases[kn] = (name, slice(h * z, (h + 1) * z))
        self.size += -(-n // ALIGN) * ALIGN

    def finalize(self):
        """Allocate params / grads and initialise (Keras glorot_normal, zeros, ones)."""
        dev = self.device
        host = torch.zeros(self.size, dtype=torch.float32)
        for name, e in self.entries.items():
            n = int(np.prod(e["shape"]))
            v = host[e["offset"]:e["offset"] + n].view(e["shape"])
            if e["fans"] is not None:
                fi, fo = e["fans"]
                std = math.sqrt(2.0 / (fi + fo)) / 0.87962566103423978
                if "pair" in e:      # two independent Keras variables, initialised one after the other
                    z = e["pair"]
                    for h in range(2):
                        t = torch.empty(e["shape"][0], z)
                        torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=self.gen)
                        v[:, h * z:(h + 1) * z] = t
                else:
                    torch.nn.init.trunc_normal_(v, 0.0, std, -2 * std, 2 * std, generator=self.gen)
            else:
                v.fill_(e["value"])
        self.flat = host.to(dev)
        self.grads = torch.zeros(self.size, dtype=torch.float32, device=dev)
        self.acc = None
        segs = np.array([s[:5] for s in self.segs], dtype=np.int64)
        chunks = [(i, st) for i, s in enumerate(self.segs) for st in range(0, s[1], CHUNK)]
        self.seg_table = torch.from_numpy(segs).to(dev)
        self.chunk_table = torch.tensor(chunks, dtype=torch.int64, device=dev)
        self.nchunk = len(chunks)
        self.nseg = len(self.segs)

    def view(self, name, grads=False):
        e = self.entries[name]
        n = int(np.prod(e["shape"]))
        src = self.grads if grads else self.flat
        return src[e["offset"]:e["offset"] + n].view(e["shape"])

    def ptr(self, name, grads=False):
        base = (self.grads if grads else self.flat).data_ptr()
        return base + 4 * self.entries[name]["offset"]

    # ---- Keras-named state dict ---------------------------------------------------------------------------
    def keras_names(self):
        out = []
        for name, e in self.entries.items():
            if "pair" in e:
                out += [k for k, (s, _) in self.aliases.items() if s == name]
            else:
                out.append(name)
        return out

    def get(self, kname, grads=False):
        if kname in self.aliases:
            s, sl = self.aliases[kname]
            v = self.view(s, grads)[:, sl]
            return v.reshape(-1) if kname.endswith("/bias") else v
        return self.view(kname, grads)

    def state_dict(self, grads=False):
        return OrderedDict((k, self.get(k, grads).detach().cpu().clone()) for k in self.keras_names())

    def load_state_dict(self, sd):
        for k in self.keras_names():
            src = sd[k]
            src = torch.as_tensor(np.asarray(src) if not torch.is_tensor(src) else src).to(torch.float32)
            self.get(k).copy_(src.to(self.device).reshape(self.get(k).shape))


class T:
    """An activation: data, gradient w.r.t. the PRE-activation of its producer, and that producer's activation."""
    __slots__ = ("data", "grad", "act")

    def __init__(self, data, grad=None, act=ACT_NONE):
        self.data, self.grad, self.act = data, grad, act

    def reshape(self, *shape):
        return T(self.data.view(*shape), None if self.grad is None else self.grad.view(*shape), self.act)


def _p(t):
    return 0 if t is None else (t if isinstance(t, int) else t.data_ptr())


# =============================================================================================================
# Ops
# =============================================================================================================
class Conv2D:
    """Conv2D / Dense (H=W=1).  fwd: mvae_conv2d_fwd; bwd: mvae_conv2d_wgrad (+ mvae_conv2d_dgrad)."""

    def __init__(self, eng, x, wname, bname, kh, kw, stride, cout, act=ACT_NONE, coord=0, need_dx=True):
        B, H, W, Cin = x.data.shape
        self.eng, self.x, self.act, self.need_dx = eng, x, act, need_dx
        self.wname = wname
        self.desc = ConvDesc(B, H, W, Cin, kh, kw, stride[0], stride[1], cout, coord, eng.precision)
        self.w, self.b = eng.ps.ptr(wname), eng.ps.ptr(bname)
        self.dw, self.db = eng.ps.ptr(wname, True), eng.ps.ptr(bname, True)
        self.y = eng.new_T((B, same_out(H, stride[0]), same_out(W, stride[1]), cout), act)
        # Dense (H = W = 1): forward / dgrad go through mvae_dense_* (tensor-core skinny GEMM, split-K scratch of this op)
        self.dense = H == 1 and W == 1 and kh == 1 and kw == 1 and coord == 0
        if self.dense:
            self.mkn = (B, Cin, cout)
            self.ws_bytes = int(eng.lib.mvae_dense_workspace_bytes(B, Cin, cout))
            self.ws = eng.empty(((self.ws_bytes + 3) // 4,)) if self.ws_bytes else None

    def fwd(self):
        L, e = self.eng.lib, self.eng
        if self.dense:
            check(L.mvae_dense_fwd(*self.mkn, _p(self.x.data), self.w, self.b, self.act, _p(self.y.data), _p(self.ws),
                                   self.ws_bytes, e.precision, e.s), "dense_fwd")
            return
        check(L.mvae_conv2d_fwd(C.byref(self.desc), _p(self.x.data), self.w, self.b, 0, 0, self.act, _p(self.y.data),
                                e.s), "conv2d_fwd")

    def bwd(self):
        L, e = self.eng.lib, self.eng
        e.wgrad(self.desc, _p(self.x.data), 0, _p(self.y.grad), self.dw, self.db)
        if self.dense and e.on_dense_wgrad is not None:
            e.on_dense_wgrad(self, e.current_level)
        if self.need_dx:
            ao = _p(self.x.data) if self.x.act != ACT_NONE else 0
            if self.dense:
                check(L.mvae_dense_dgrad(*self.mkn, _p(self.y.grad), self.w, ao, self.x.act, _p(self.x.grad), _p(self.ws),
                                         self.ws_bytes, e.precision, e.s), "dense_dgrad")
                return
            check(L.mvae_conv2d_dgrad(C.byref(self.desc), _p(self.y.grad), self.w, 0, 0, ao, self.x.act,
                                      _p(self.x.grad), e.s), "conv2d_dgrad")


class Conv2DTranspose:
    """Conv2DTranspose == dgrad of the SAME forward conv that maps the output back to the input."""

    def __init__(self, eng, x, wname, bname, kh, kw, stride, cout):
        B, H, W, Cin = x.data.shape
        assert x.act == ACT_NONE
        self.eng, self.x = eng, x
        Ho, Wo = H * stride[0], W * stride[1]
        # forward conv of the pair: (B,Ho,Wo,cout) -> (B,H,W,Cin)
        self.desc = ConvDesc(B, Ho, Wo, cout, kh, kw, stride[0], stride[1], Cin, 0, eng.precision)
        self.w, self.b = eng.ps.ptr(wname), eng.ps.ptr(bname)
        self.dw, self.db = eng.ps.ptr(wname, True), eng.ps.ptr(bname, True)
        self.y = eng.new_T((B, Ho, Wo, cout), ACT_NONE)
        self.M, self.cout = B * Ho * Wo, cout

    def fwd(self):
        L, e = self.eng.lib, self.eng
        check(L.mvae_conv2d_dgrad(C.byref(self.desc), _p(self.x.data), self.w, self.b, 0, 0, ACT_NONE, _p(self.y.data),
                                  e.s), "conv2d_transpose_fwd")

    def bwd(self):
        L, e = self.eng.lib, self.eng

        e.wgrad(self.desc, _p(self.y.grad), 0, _p(self.x.data), self.dw, 0)
        e.side(lambda: check(L.mvae_colsum(_p(self.y.grad), self.db, self.M, self.cout, e.s), "colsum"))
        check(L.mvae_conv2d_fwd(C.byref(self.desc), _p(self.y.grad), self.w, 0, 0, 0, ACT_NONE, _p(self.x.grad), e.s),
              "conv2d_transpose_dgrad")


class MobileNetV3:
    """layer_blocks.py:556-648: 1x1+ReLU -> depthwise 3x3+ReLU -> squeeze-excite(BN) -> 1x1 -> + input."""

    def __init__(self, eng, x, prefix, filters):
        B, H, W, Cin = x.data.shape
        ps = eng.ps
        self.eng, self.x, self.B, self.H, self.W, self.F = eng, x, B, H, W, filters
        self.d0 = ConvDesc(B, H, W, Cin, 1, 1, 1, 1, filters, 0, eng.precision)
        self.d2 = ConvDesc(B, H, W, filters, 1, 1, 1, 1, Cin, 0, eng.precision)
        n = lambda s: prefix + s
        self.pn = dict(w0=n("conv0/kernel"), b0=n("conv0/bias"), wd=n("conv1/depthwise_kernel"), bd=n("conv1/bias"),
                       s0=n("squeeze_excite_dense0/kernel"), sb0=n("squeeze_excite_dense0/bias"),
                       g=n("squeeze_excite_batchnorm0/gamma"), be=n("squeeze_excite_batchnorm0/beta"),
                       mm=n("squeeze_excite_batchnorm0/moving_mean"), mv=n("squeeze_excite_batchnorm0/moving_variance"),
                       s1=n("squeeze_excite_dense1/kernel"), sb1=n("squeeze_excite_dense1/bias"),
                       w2=n("conv2/kernel"), b2=n("conv2/bias"))
        self.P = {k: ps.ptr(v) for k, v in self.pn.items()}
        self.G = {k: ps.ptr(v, True) for k, v in self.pn.items()}
        self.a = eng.empty((B, H, W, filters))
        self.u = eng.empty((B, H, W, filters))
        self.gate = eng.empty((B, filters))
        self.ws = eng.empty((eng.lib.mvae_se_gate_ws_floats(B, filters),))
        self.gap = eng.zeros(B * filters)
        self.se_stat = eng.zeros(128)                 # 64 doubles: batch sums of the folded squeeze-excite gate
        self.y = eng.new_T((B, H, W, Cin), ACT_NONE)
        if eng.training:
            self.dv = eng.empty((B, H, W, filters))
            self.da = eng.empty((B, H, W, filters))
            self.dg = eng.zeros(B * filters, bwd=True)
            self.se_bstat = eng.zeros(128, bwd=True)
            self.dgap = eng.empty((B, filters))

    def fwd(self):
        L, e, P = self.eng.lib, self.eng, self.P
        check(L.mvae_conv2d_fwd(C.byref(self.d0), _p(self.x.data), P["w0"], P["b0"], 0, 0, ACT_RELU, _p(self.a), e.s),
              "mbv3 conv0")
        check(L.mvae_dwconv3x3_fwd(_p(self.a), P["wd"], P["bd"], _p(self.u), self.gap.ptr, self.B, self.H, self.W, self.F,
                                   e.s), "mbv3 dw")
        check(L.mvae_se_gate_fwd(self.gap.ptr, P["s0"], P["sb0"], P["g"], P["be"], P["s1"], P["sb1"], P["mm"], P["mv"],
                                 _p(self.gate), _p(self.ws), self.B, self.F, self.H * self.W, SE_BN_EPS, SE_BN_MOM,
                                 1 if e.training else 0, e.s), "mbv3 se")
        check(L.mvae_conv2d_fwd(C.byref(self.d2), _p(self.u), P["w2"], P["b2"], _p(self.gate), _p(self.x.data), ACT_NONE,
                                _p(self.y.data), e.s), "mbv3 conv2")

    def bwd(self):
        L, e, P, G = self.eng.lib, self.eng, self.P, self.G
        dy = _p(self.y.grad)
        check(L.mvae_conv2d_dgrad(C.byref(self.d2), dy, P["w2"], 0, 0, 0, ACT_NONE, _p(self.dv), e.s), "mbv3 conv2 dgrad")
        e.wgrad(self.d2, _p(self.u), _p(self.gate), dy, G["w2"], G["b2"])
        check(L.mvae_se_dgate_reduce(_p(self.dv), _p(self.u), self.dg.ptr, self.B, self.H * self.W, self.F, e.s),
              "mbv3 dgate")
        check(L.mvae_se_gate_bwd(self.dg.ptr, P["s0"], P["g"], P["be"], P["s1"], _p(self.ws), _p(self.dgap), G["s0"], G["sb0"],
                                 G["g"], G["be"], G["s1"], G["sb1"], self.B, self.F, self.H * self.W, e.s), "mbv3 se bwd")
        check(L.mvae_dwconv3x3_bwd(_p(self.a), _p(self.u), _p(self.dv), _p(self.gate), _p(self.dgap), P["wd"],
                                   _p(self.da), G["wd"], G["bd"], self.B, self.H, self.W, self.F, e.s), "mbv3 dw bwd")
        e.wgrad(self.d0, _p(self.x.data), 0, _p(self.da), G["w0"], G["b0"])
        if self.x.grad is not None:
            ao = _p(self.x.data) if self.x.act != ACT_NONE else 0
            check(L.mvae_conv2d_dgrad(C.byref(self.d0), _p(self.da), P["w0"], 0, dy, ao, self.x.act, _p(self.x.grad),
                                      e.s), "mbv3 conv0 dgrad")


class FusedMBV3Chain:
    """A run of mobilenetV3 blocks on one image size, on the fused tile kernels (csrc/mbv3_fused.cu): n blocks take n+1
    launches forward and n+1 backward -- [conv2 + residual of block j-1 | conv0 + depthwise + GAP of block j] and
    [depthwise^T + conv0 dgrad + residual of block j+1 | gate-gradient sums of block j] -- with the squeeze-excite gate
    kernels (the batch-wide BatchNorm, layer_blocks.py:447-449) in between.  The 1x1 weight gradients stay deferred
    mvae_conv2d_wgrad calls.  TF32 only: with MVAE_PREC_FP32 (or MVAE_NO_FUSED_MBV3=1) the blocks run layer by layer."""

    def __init__(self, eng, blocks):
        self.eng, self.blocks = eng, blocks
        self.y = blocks[-1].y
        b0 = blocks[0]
        self.dims = (b0.B, b0.H, b0.W, b0.F)

    @staticmethod
    def supported(eng, blk):
        B, H, W, Cin = blk.x.data.shape
        return (blk.x.act == ACT_NONE and (blk.x.grad is not None or not eng.training)
                and bool(eng.lib.mvae_mbv3_fused_supported(B, H, W, Cin, blk.F)))

    def fused(self):
        return self.eng.precision == PREC_TF32 and self.eng.fuse_mbv3

    def fold(self, direction):
        """Squeeze-excite gate inside the tile kernels: whole-image tiles only (a CTA then owns complete images).
        eng.fold_se: "" (off), "fwd", "bwd" or "both"."""
        return self.eng.fold_se in (direction, "both") and self.dims[1] * self.dims[2] <= 256

    def _se_fwd(self, m):
        L, e, P = self.eng.lib, self.eng, m.P
        check(L.mvae_se_gate_fwd(m.gap.ptr, P["s0"], P["sb0"], P["g"], P["be"], P["s1"], P["sb1"], P["mm"], P["mv"],
                                 _p(m.gate), _p(m.ws), m.B, m.F, m.H * m.W, SE_BN_EPS, SE_BN_MOM,
                                 1 if e.training else 0, e.s), "mbv3 se")

    def _se_bwd(self, m):
        L, e, P, G = self.eng.lib, self.eng, m.P, m.G
        check(L.mvae_se_gate_bwd(m.dg.ptr, P["s0"], P["g"], P["be"], P["s1"], _p(m.ws), _p(m.dgap), G["s0"], G["sb0"],
                                 G["g"], G["be"], G["s1"], G["sb1"], m.B, m.F, m.H * m.W, e.s), "mbv3 se bwd")

    def fill_fwd(self, a, j, fold):
        """Arguments of forward launch j (0..n) of the chain: [conv2 + residual of block j-1 | conv0 + depthwise of block j]."""
        e, bl, n = self.eng, self.blocks, len(self.blocks)
        a.B, a.H, a.W, a.C = self.dims
        if j > 0:
            m = bl[j - 1]
            a.u_prev, a.x_prev, a.y = _p(m.u), _p(m.x.data), _p(m.y.data)
            a.w2, a.b2 = m.P["w2"], m.P["b2"]
            if fold:
                P = m.P
                a.se_gamma_prev, a.se_beta_prev, a.se_w1_prev, a.se_b1_prev = P["g"], P["be"], P["s1"], P["sb1"]
                a.se_mm_prev, a.se_mv_prev, a.se_ws_prev, a.gate_out_prev = P["mm"], P["mv"], _p(m.ws), _p(m.gate)
                a.se_stat_prev = m.se_stat.ptr
                a.bn_eps, a.bn_momentum, a.training = SE_BN_EPS, SE_BN_MOM, 1 if e.training else 0
            else:
                a.gate_prev = _p(m.gate)
        if j < n:
            m = bl[j]
            a.x = _p(m.x.data) if j == 0 else None
            a.w0, a.b0, a.wd, a.bd = m.P["w0"], m.P["b0"], m.P["wd"], m.P["bd"]
            a.a = _p(m.a) if e.training else None
            a.u = _p(m.u)
            if fold:
                a.se_w0, a.se_b0, a.se_ws, a.se_stat = m.P["s0"], m.P["sb0"], _p(m.ws), m.se_stat.ptr
            else:
                a.gap_sum = m.gap.ptr

    def fill_bwd(self, a, k, fold):
        """Arguments of backward launch k (n..0): [depthwise^T + conv0 dgrad of block k | gate-gradient sums of block k-1]."""
        bl, n = self.blocks, len(self.blocks)
        a.B, a.H, a.W, a.C = self.dims
        if k < n:
            m = bl[k]
            a.dy, a.u, a.a, a.gate = _p(m.y.grad), _p(m.u), _p(m.a), _p(m.gate)
            a.w2, a.wd, a.w0 = m.P["w2"], m.P["wd"], m.P["w0"]
            a.da, a.dx, a.dwd, a.dbd = _p(m.da), _p(m.x.grad), m.G["wd"], m.G["bd"]
            if fold:
                a.se_w0, a.se_gamma, a.se_ws, a.se_bstat = m.P["s0"], m.P["g"], _p(m.ws), m.se_bstat.ptr
            else:
                a.dgap = _p(m.dgap)
        if k > 0:
            m = bl[k - 1]
            if k == n:
                a.dy = _p(m.y.grad)
            a.w2_prev, a.u_prev, a.dgate_prev = m.P["w2"], _p(m.u), m.dg.ptr
            if fold:
                a.se_w1_prev, a.se_ws_prev, a.se_bstat_prev = m.P["s1"], _p(m.ws), m.se_bstat.ptr

    def defer_wgrads(self, k):
        m, e = self.blocks[k], self.eng
        e.wgrad(m.d2, _p(m.u), _p(m.gate), _p(m.y.grad), m.G["w2"], m.G["b2"])
        e.wgrad(m.d0, _p(m.x.data), 0, _p(m.da), m.G["w0"], m.G["b0"])

    def fwd(self):
        if not self.fused():
            for m in self.blocks:
                m.fwd()
            return
        L, e, bl, n = self.eng.lib, self.eng, self.blocks, len(self.blocks)
        fold = self.fold("fwd")
        for j in range(n + 1):
            a = Mbv3FwdArgs()
            self.fill_fwd(a, j, fold)
            check(L.mvae_mbv3_fused_fwd(C.byref(a), e.s), "mbv3 fused fwd")
            if j < n and not fold:
                self._se_fwd(bl[j])

    def bwd(self):
        if not self.fused():
            for m in reversed(self.blocks):
                m.bwd()
            return
        L, e, bl, n = self.eng.lib, self.eng, self.blocks, len(self.blocks)
        fold = self.fold("bwd")
        for k in range(n, -1, -1):
            a = Mbv3BwdArgs()
            self.fill_bwd(a, k, fold)
            check(L.mvae_mbv3_fused_bwd(C.byref(a), e.s), "mbv3 fused bwd")
            if k < n:
                self.defer_wgrads(k)
            if k > 0:
                if fold:
                    # the gate's dgap comes out of the next launch; only the squeeze-excite WEIGHT gradients are left, and
                    # nothing in the chain waits for them
                    e.side(lambda m=bl[k - 1]: self._se_bwd(m), lane=9)
                else:
                    self._se_bwd(bl[k - 1])


class BatchedFusedChain:
    """The same chain of every coarse pyramid level in ONE sequence of launches: step j of all members goes out as a
    single mvae_mbv3_fused_*_batched call (each member with its own image size, weights and buffers) and the squeeze-excite
    gates as one mvae_se_gate_*_batched call.  Members must be FusedMBV3Chain objects with the same number of blocks."""

    def __init__(self, eng, chains):
        self.eng, self.chains, self.n, self.nblk = eng, chains, len(chains), len(chains[0].blocks)

    def _se(self, j, fwd):
        L, e = self.eng.lib, self.eng
        ms = [c.blocks[j] for c in self.chains]
        m0 = ms[0]
        A = lambda f: _pa([f(m) for m in ms])
        HW = _ia([m.H * m.W for m in ms])
        if fwd:
            check(L.mvae_se_gate_fwd_batched(self.n, A(lambda m: m.gap.ptr), A(lambda m: m.P["s0"]), A(lambda m: m.P["sb0"]),
                                             A(lambda m: m.P["g"]), A(lambda m: m.P["be"]), A(lambda m: m.P["s1"]),
                                             A(lambda m: m.P["sb1"]), A(lambda m: m.P["mm"]), A(lambda m: m.P["mv"]),
                                             A(lambda m: _p(m.gate)), A(lambda m: _p(m.ws)), m0.B, m0.F, HW, SE_BN_EPS,
                                             SE_BN_MOM, 1 if e.training else 0, e.s), "mbv3 se (levels)")
        else:
            check(L.mvae_se_gate_bwd_batched(self.n, A(lambda m: m.dg.ptr), A(lambda m: m.P["s0"]), A(lambda m: m.P["g"]),
                                             A(lambda m: m.P["be"]), A(lambda m: m.P["s1"]), A(lambda m: _p(m.ws)),
                                             A(lambda m: _p(m.dgap)), A(lambda m: m.G["s0"]), A(lambda m: m.G["sb0"]),
                                             A(lambda m: m.G["g"]), A(lambda m: m.G["be"]), A(lambda m: m.G["s1"]),
                                             A(lambda m: m.G["sb1"]), m0.B, m0.F, HW, e.s), "mbv3 se bwd (levels)")

    def fwd(self):
        if not self.chains[0].fused():
            for c in self.chains:
                c.fwd()
            return
        L, e = self.eng.lib, self.eng
        for j in range(self.nblk + 1):
            arr = (Mbv3FwdArgs * self.n)()
            for l, c in enumerate(self.chains):
                c.fill_fwd(arr[l], j, False)
            check(L.mvae_mbv3_fused_fwd_batched(self.n, arr, e.s), "mbv3 fused fwd (levels)")
            if j < self.nblk:
                self._se(j, True)

    def bwd(self):
        if not self.chains[0].fused():
            for c in self.chains:
                c.bwd()
            return
        L, e = self.eng.lib, self.eng
        for k in range(self.nblk, -1, -1):
            arr = (Mbv3BwdArgs * self.n)()
            for l, c in enumerate(self.chains):
                c.fill_bwd(arr[l], k, False)
            check(L.mvae_mbv3_fused_bwd_batched(self.n, arr, e.s), "mbv3 fused bwd (levels)")
            if k < self.nblk:
                for c in self.chains:
                    c.defer_wgrads(k)
            if k > 0:
                self._se(k - 1, False)


def fuse_chains(eng, ops):
    """Replace every run of consecutive fusable mobilenetV3 blocks of equal shape by one FusedMBV3Chain."""
    out, run = [], []

    def close():
        if run:
            out.append(FusedMBV3Chain(eng, list(run)))
            run.clear()

    for op in ops:
        if isinstance(op, MobileNetV3) and FusedMBV3Chain.supported(eng, op) and \
                (not run or (run[-1].H, run[-1].W, run[-1].F) == (op.H, op.W, op.F)):
            run.append(op)
            continue
        close()
        if isinstance(op, MobileNetV3) and FusedMBV3Chain.supported(eng, op):
            run.append(op)
        else:
            out.append(op)
    close()
    return out


class Reparam:
    """sample Lambda (multiscale_vae.py:372-383) fused with the per-scale KL (:485-488)."""

    def __init__(self, eng, mulv, eps, kl, zdim):
        self.eng, self.mulv, self.eps, self.kl, self.z = eng, mulv, eps, kl, zdim
        self.B = mulv.data.shape[0]
        self.y = eng.new_T((self.B, 1, 1, zdim), ACT_NONE)

    def fwd(self):
        e = self.eng
        check(e.lib.mvae_reparam_kl_fwd(_p(self.mulv.data), _p(self.eps), _p(self.y.data), _p(self.kl), self.B, self.z,
                                        e.logvar_scale, e.sample_std, e.s), "reparam_kl_fwd")

    def bwd(self):
        e = self.eng
        check(e.lib.mvae_reparam_kl_bwd(_p(self.mulv.data), _p(self.eps), _p(self.y.grad), _p(self.mulv.grad), self.B,
                                        self.z, e.logvar_scale, e.sample_std, e.kl_factor / e.B, e.s), "reparam_kl_bwd")


class Tail:
    """BatchNormalization(.999, 1e-4) + Conv2D 1x1 -> C (multiscale_vae.py:420-431)."""

    def __init__(self, eng, x, prefix, cout):
        B, H, W, F = x.data.shape
        assert x.act == ACT_NONE
        ps = eng.ps
        self.eng, self.x, self.M, self.F, self.Co = eng, x, B * H * W, F, cout
        self.pn = dict(g=prefix + "batchnorm/gamma", be=prefix + "batchnorm/beta", mm=prefix + "batchnorm/moving_mean",
                       mv=prefix + "batchnorm/moving_variance", w=prefix + "conv_out/kernel", b=prefix + "conv_out/bias")
        self.P = {k: ps.ptr(v) for k, v in self.pn.items()}
        self.G = {k: ps.ptr(v, True) for k, v in self.pn.items()}
        self.sums = eng.zeros(4 * F)          # 2F doubles
        self.stats = eng.empty((2 * F,))
        self.y = eng.new_T((B, H, W, cout), ACT_NONE)
        if eng.training:
            self.red = eng.zeros(F * cout + cout, bwd=True)

    def fwd(self):
        L, e, P = self.eng.lib, self.eng, self.P
        if e.training:
            check(L.mvae_bn_stats(_p(self.x.data), self.sums.ptr, self.M, self.F, e.s), "bn_stats")
        check(L.mvae_bn_convout_fwd(_p(self.x.data), self.sums.ptr, P["g"], P["be"], P["mm"], P["mv"], P["w"], P["b"],
                                    _p(self.y.data), _p(self.stats), self.M, self.F, self.Co, DEC_BN_EPS, DEC_BN_MOM,
                                    1 if e.training else 0, e.s), "bn_convout_fwd")

    def bwd(self):
        L, e, P, G = self.eng.lib, self.eng, self.P, self.G
        check(L.mvae_bn_convout_bwd(_p(self.x.data), _p(self.y.grad), _p(self.stats), P["g"], P["be"], P["w"],
                                    self.red.ptr, _p(self.x.grad), G["g"], G["be"], G["w"], G["b"], self.M, self.F,
                                    self.Co, e.s), "bn_convout_bwd")


# =============================================================================================================
# Level-batched groups: layer k of every pyramid level in one C-ABI call (one launch when the kernels allow it)
# =============================================================================================================
def _pa(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(p) if p else None for p in ptrs])


def _ia(vals):
    return (C.c_int * len(vals))(*vals)


def _descs(ds):
    arr = (ConvDesc * len(ds))()
    for i, d in enumerate(ds):
        C.memmove(C.byref(arr[i]), C.byref(d), C.sizeof(ConvDesc))
    return arr


class BatchedMobileNetV3:
    def __init__(self, eng, blocks):
        self.eng, self.b, self.n = eng, blocks, len(blocks)
        b0 = blocks[0]
        self.B, self.F = b0.B, b0.F
        self.H, self.W, self.HW = _ia([m.H for m in blocks]), _ia([m.W for m in blocks]), _ia([m.H * m.W for m in blocks])
        A = lambda f: _pa([f(m) for m in blocks])
        self.x, self.a, self.u, self.gate = A(lambda m: _p(m.x.data)), A(lambda m: _p(m.a)), A(lambda m: _p(m.u)), A(lambda m: _p(m.gate))
        self.ws, self.gap, self.y = A(lambda m: _p(m.ws)), None, A(lambda m: _p(m.y.data))
        self.P = {k: A(lambda m, k=k: m.P[k]) for k in b0.P}
        self.G = {k: A(lambda m, k=k: m.G[k]) for k in b0.G}
        self.train = eng.training
        if self.train:
            self.dy, self.dv, self.da = A(lambda m: _p(m.y.grad)), A(lambda m: _p(m.dv)), A(lambda m: _p(m.da))
            self.dgap = A(lambda m: _p(m.dgap))
            self.has_dx = b0.x.grad is not None
            self.dx = A(lambda m: _p(m.x.grad)) if self.has_dx else None
            self.xact = b0.x.act
            self.ao = A(lambda m: _p(m.x.data)) if self.xact != ACT_NONE else None

    def _late(self):
        # the zero-arena slices get their addresses after the engine is built
        if self.gap is None:
            self.gap = _pa([m.gap.ptr for m in self.b])
            if self.train:
                self.dg = _pa([m.dg.ptr for m in self.b])

    def fwd(self):
        L, e, P, n = self.eng.lib, self.eng, self.P, self.n
        self._late()
        d0, d2 = _descs([m.d0 for m in self.b]), _descs([m.d2 for m in self.b])
        check(L.mvae_conv2d_fwd_batched(n, d0, self.x, P["w0"], P["b0"], None, None, ACT_RELU, self.a, e.s), "mbv3 conv0")
        check(L.mvae_dwconv3x3_fwd_batched(n, self.a, P["wd"], P["bd"], self.u, self.gap, self.B, self.H, self.W, self.F, e.s),
              "mbv3 dw")
        check(L.mvae_se_gate_fwd_batched(n, self.gap, P["s0"], P["sb0"], P["g"], P["be"], P["s1"], P["sb1"], P["mm"], P["mv"],
                                         self.gate, self.ws, self.B, self.F, self.HW, SE_BN_EPS, SE_BN_MOM,
                                         1 if e.training else 0, e.s), "mbv3 se")
        check(L.mvae_conv2d_fwd_batched(n, d2, self.u, P["w2"], P["b2"], self.gate, self.x, ACT_NONE, self.y, e.s), "mbv3 conv2")

    def bwd(self):
        L, e, P, G, n = self.eng.lib, self.eng, self.P, self.G, self.n
        self._late()
        d0, d2 = _descs([m.d0 for m in self.b]), _descs([m.d2 for m in self.b])
        check(L.mvae_conv2d_dgrad_batched(n, d2, self.dy, P["w2"], None, None, None, ACT_NONE, self.dv, e.s), "mbv3 conv2 dgrad")
        e.side(lambda: check(L.mvae_conv2d_wgrad_batched(n, d2, self.u, self.gate, self.dy, G["w2"], G["b2"], e.s),
                             "mbv3 conv2 wgrad"))
        check(L.mvae_se_dgate_reduce_batched(n, self.dv, self.u, self.dg, self.B, self.HW, self.F, e.s), "mbv3 dgate")
        check(L.mvae_se_gate_bwd_batched(n, self.dg, P["s0"], P["g"], P["be"], P["s1"], self.ws, self.dgap, G["s0"], G["sb0"],
                                         G["g"], G["be"], G["s1"], G["sb1"], self.B, self.F, self.HW, e.s), "mbv3 se bwd")
        check(L.mvae_dwconv3x3_bwd_batched(n, self.a, self.u, self.dv, self.gate, self.dgap, P["wd"], self.da, G["wd"], G["bd"],
                                           self.B, self.H, self.W, self.F, e.s), "mbv3 dw bwd")
        e.side(lambda: check(L.mvae_conv2d_wgrad_batched(n, d0, self.x, None, self.da, G["w0"], G["b0"], e.s),
                             "mbv3 conv0 wgrad"))
        if self.has_dx:
            check(L.mvae_conv2d_dgrad_batched(n, d0, self.da, P["w0"], None, self.dy, self.ao, self.xact, self.dx, e.s),
                  "mbv3 conv0 dgrad")


class BatchedConv2D:
    def __init__(self, eng, convs):
        self.eng, self.c, self.n = eng, convs, len(convs)
        c0 = convs[0]
        A = lambda f: _pa([f(m) for m in convs])
        self.x, self.y = A(lambda m: _p(m.x.data)), A(lambda m: _p(m.y.data))
        self.w, self.b, self.dw, self.db = A(lambda m: m.w), A(lambda m: m.b), A(lambda m: m.dw), A(lambda m: m.db)
        self.act, self.need_dx = c0.act, c0.need_dx
        if eng.training:
            self.dy = A(lambda m: _p(m.y.grad))
            if self.need_dx:
                self.dx = A(lambda m: _p(m.x.grad))
                self.xact = c0.x.act
                self.ao = A(lambda m: _p(m.x.data)) if self.xact != ACT_NONE else None

    def fwd(self):
        e = self.eng
        check(e.lib.mvae_conv2d_fwd_batched(self.n, _descs([m.desc for m in self.c]), self.x, self.w, self.b, None, None,
                                            self.act, self.y, e.s), "conv2d_fwd")

    def bwd(self):
        L, e = self.eng.lib, self.eng
        d = _descs([m.desc for m in self.c])
        e.side(lambda: check(L.mvae_conv2d_wgrad_batched(self.n, d, self.x, None, self.dy, self.dw, self.db, e.s), "conv2d_wgrad"))
        if self.need_dx:
            check(L.mvae_conv2d_dgrad_batched(self.n, d, self.dy, self.w, None, None, self.ao, self.xact, self.dx, e.s),
                  "conv2d_dgrad")


class BatchedConv2DTranspose:
    def __init__(self, eng, convs):
        self.eng, self.c, self.n = eng, convs, len(convs)
        A = lambda f: _pa([f(m) for m in convs])
        self.x, self.y = A(lambda m: _p(m.x.data)), A(lambda m: _p(m.y.data))
        self.w, self.b, self.dw = A(lambda m: m.w), A(lambda m: m.b), A(lambda m: m.dw)
        if eng.training:
            self.dy, self.dx = A(lambda m: _p(m.y.grad)), A(lambda m: _p(m.x.grad))

    def fwd(self):
        e = self.eng
        check(e.lib.mvae_conv2d_dgrad_batched(self.n, _descs([m.desc for m in self.c]), self.x, self.w, self.b, None, None,
                                              ACT_NONE, self.y, e.s), "conv2d_transpose_fwd")

    def bwd(self):
        L, e = self.eng.lib, self.eng
        d = _descs([m.desc for m in self.c])

        def wgrad():
            check(L.mvae_conv2d_wgrad_batched(self.n, d, self.dy, None, self.x, self.dw, None, e.s), "conv2d_transpose_wgrad")
            for m in self.c:
                check(L.mvae_colsum(_p(m.y.grad), m.db, m.M, m.cout, e.s), "colsum")

        e.side(wgrad)
        check(L.mvae_conv2d_fwd_batched(self.n, d, self.dy, self.w, None, None, None, ACT_NONE, self.dx, e.s),
              "conv2d_transpose_dgrad")


def _same_layer(descs):
    d0 = descs[0]
    key = lambda d: (d.Cin, d.Cout, d.kh, d.kw, d.sh, d.sw, d.coord_mode)
    return all(key(d) == key(d0) for d in descs) and d0.coord_mode == 0 and d0.Cin % 32 == 0 and d0.Cout % 32 == 0


def make_batched(eng, group):
    """group: the op at one position of every level's list.  Returns a batched op or None."""
    if len(group) < 2 or len(group) > 8:
        return None
    t = type(group[0])
    if any(type(g) is not t for g in group):
        return None
    if t is FusedMBV3Chain:
        if len({len(g.blocks) for g in group}) == 1 and len({g.dims[0] for g in group}) == 1:
            return BatchedFusedChain(eng, group)
    elif t is MobileNetV3:
        if _same_layer([g.d0 for g in group]) and group[0].F % 4 == 0:
            return BatchedMobileNetV3(eng, group)
    elif t is Conv2D:
        if _same_layer([g.desc for g in group]):
            return BatchedConv2D(eng, group)
    elif t is Conv2DTranspose:
        if _same_layer([g.desc for g in group]):
            return BatchedConv2DTranspose(eng, group)
    return None


# =============================================================================================================
# Variable declarations (Keras shapes / initialiser fans / regularisers)
# =============================================================================================================
def declare_conv(ps, name, kh, kw, cin, cout, reg, transpose=False):
    shape = (kh, kw, cout, cin) if transpose else (kh, kw, cin, cout)
    ps.add(name + "/kernel", shape, reg, fans=(kh * kw * shape[2], kh * kw * shape[3]))
    ps.add(name + "/bias", (cout,))


def declare_dense(ps, name, kin, kout, reg):
    ps.add(name + "/kernel", (kin, kout), reg, fans=(kin, kout))
    ps.add(name + "/bias", (kout,))


def declare_bn(ps, name, c):
    ps.add(name + "/gamma", (c,), value=1.0)
    ps.add(name + "/beta", (c,))
    ps.add(name + "/moving_mean", (c,), trainable=False)
    ps.add(name + "/moving_variance", (c,), value=1.0, trainable=False)


def declare_mbv3(ps, prefix, cin, f):
    """layer_blocks.py:556-648 + squeeze_excite_block(use_batchnorm=True), :418-462"""
    declare_conv(ps, prefix + "conv0", 1, 1, cin, f, REG_L1)
    ps.add(prefix + "conv1/depthwise_kernel", (3, 3, f, 1), REG_L1, fans=(9 * f, 9))
    ps.add(prefix + "conv1/bias", (f,))
    declare_dense(ps, prefix + "squeeze_excite_dense0", f, f, REG_L1)
    declare_bn(ps, prefix + "squeeze_excite_batchnorm0", f)
    declare_dense(ps, prefix + "squeeze_excite_dense1", f, f, REG_L1)
    declare_conv(ps, prefix + "conv2", 1, 1, f, cin, REG_L1)


class _Z:
    """A slice of the per-step zero arena (resolved once the arena is allocated)."""
    __slots__ = ("offset", "n", "ptr")

    def __init__(self, offset, n):
        self.offset, self.n, self.ptr = offset, n, 0


# =============================================================================================================
# Model description (shared by every Engine of a model)
# =============================================================================================================
class Spec:
    def __init__(self, input_dims, z_dims, encoder, decoder, v0, v1, sample_std, coord_conv, logvar_scale, diff_mode):
        H, W, Cc = input_dims
        self.H, self.W, self.C = H, W, Cc
        self.z_dims, self.levels = list(z_dims), len(z_dims)
        self.enc, self.dec = encoder, decoder
        self.v0, self.v1, self.sample_std = float(v0), float(v1), float(sample_std)
        self.coord = {None: 0, "xy": 2, "xyr": 3}[coord_conv]
        self.logvar_scale = float(logvar_scale)
        self.diff_mode = {"no_upsample": DIFF_NO_UPSAMPLE, "laplacian": DIFF_LAPLACIAN}[diff_mode]
        self.conv_base_filters = 32                                   # multiscale_vae.py:50
        self.taps = gaussian_kernel((3, 3), (2, 2)).astype(np.float32)  # multiscale_vae.py:56-57
        if self.levels < 2:
            raise ValueError("MultiscaleVAE needs at least 2 levels (the reference merge loop, "
                             "multiscale_vae.py:210-222, leaves its output unbound for 1 level)")
        if H % (1 << (self.levels - 1)) or W % (1 << (self.levels - 1)):
            raise ValueError(f"input {H}x{W} is not divisible by 2^(levels-1) = {1 << (self.levels - 1)}: the "
                             "reference graph is inconsistent for odd scales (pool ceil vs int(x/2), "
                             "multiscale_vae.py:123,308-311)")
        self.scales = [(H >> i, W >> i, Cc) for i in range(self.levels)]   # multiscale_vae.py:111-127

    @staticmethod
    def entries(cfg):
        f, k, s = cfg["filters"], cfg["kernel_size"], cfg["strides"]
        if len(f) != len(k) or len(f) != len(s) or len(f) <= 0:          # layer_blocks.py:918-924
            raise ValueError("len(filters) [{0}] should be equal to len(kernel_size) [{1}] and len(strides) [{2}]"
                             .format(len(f), len(k), len(s)))
        return [(int(a), (int(b[0]), int(b[1])), (int(c[0]), int(c[1]))) for a, b, c in zip(f, k, s)]

    def declare_params(self, ps):
        """Create every variable in Keras creation order (encoders then decoders, multiscale_vae.py:172-200)."""
        self.before_flatten = []

        conv = lambda *a, **k: declare_conv(ps, *a, **k)
        dense = lambda *a, **k: declare_dense(ps, *a, **k)
        bn = lambda *a, **k: declare_bn(ps, *a, **k)
        mbv3 = lambda *a, **k: declare_mbv3(ps, *a, **k)

        for i in range(self.levels):
            h, w, c = self.scales[i]
            p = f"encoder_{i}_"
            conv(p + "conv_base", 3, 3, c + self.coord, self.conv_base_filters, REG_L2)
            prev = self.conv_base_filters
            for j, (f, k, s) in enumerate(self.entries(self.enc)):
                if s != (1, 1) or f != prev:
                    conv(f"{p}_{j}_conv", k[0], k[1], prev, f, REG_L1)
                    h, w = same_out(h, s[0]), same_out(w, s[1])
                mbv3(f"{p}_{j}_mobilenetV3_", f, f)
                prev = f
            self.before_flatten.append((h, w, prev))
            K, z = h * w * prev, self.z_dims[i]
            ps.add_fused_pair(p + "mu_log_var/kernel", K, z, REG_L2, (p + "mu/kernel", p + "log_var/kernel"), (K, z))
            ps.add_fused_pair(p + "mu_log_var/bias", 1, z, REG_NONE, (p + "mu/bias", p + "log_var/bias"), None)
        for i in range(self.levels):
            h, w, c = self.before_flatten[i]
            p = f"decoder_{i}_"
            dense(p + "dense", self.z_dims[i], h * w * c, REG_L2)
            prev = c
            for j, (f, k, s) in enumerate(self.entries(self.dec)):
                if s != (1, 1) or f != prev:
                    conv(f"{p}_{j}_conv_transpose", k[0], k[1], prev, f, REG_L1, transpose=True)
                    h, w = h * s[0], w * s[1]
                mbv3(f"{p}_{j}_mobilenetV3_", f, f)
                prev = f
            if (h, w) != self.scales[i][:2]:
                raise ValueError(f"level {i}: decoder emits {h}x{w} but the scale is {self.scales[i][0]}x"
                                 f"{self.scales[i][1]}: the total stride must divide every scale")
            bn(p + "batchnorm", prev)
            conv(p + "conv_out", 1, 1, prev, self.scales[i][2], REG_L2)


# =============================================================================================================
# Engine
# =============================================================================================================
class Engine:
    def __init__(self, spec: Spec, ps: ParamStore, B, training, precision=PREC_FP32, corrupt=False):
        self.spec, self.ps, self.B, self.training, self.precision = spec, ps, int(B), training, precision
        self.corrupt = bool(corrupt)          # the split reads x_in (noise / channel dropout applied), the loss the clean x
        self.lib = _lib.load()
        self.device = ps.device
        self.s = 0
        self.sample_std, self.logvar_scale = spec.sample_std, spec.logvar_scale
        self.r_factor, self.kl_factor = 1.0, 1.0
        self._zreq, self._zsize = [[], []], [0, 0]
        self._convs = []
        # fused mobilenetV3 tile kernels (TF32 only; MVAE_NO_FUSED_MBV3=1 keeps the layer-by-layer launches)
        self.fuse_mbv3 = os.environ.get("MVAE_NO_FUSED_MBV3") != "1"
        # squeeze-excite gate folded into the fused launches: measured on cfg2 (B200) the separate gate kernels are as fast
        # (1.39 vs 1.44 ms/step) -- every folded piece is a chain of dependent global round trips -- so it is opt-in
        self.fold_se = os.environ.get("MVAE_FOLD_SE", "")
        self._build()
        # per-step accumulators live in two arenas, each cleared by one memset: [0] forward (GAP sums, BN sums, loss
        # sums, optimiser norms), [1] backward (gate-gradient sums, tail reductions)
        self.arena = torch.zeros(max(self._zsize[0], ALIGN), dtype=torch.float32, device=self.device)
        self.arena_bwd = torch.zeros(max(self._zsize[1], ALIGN), dtype=torch.float32, device=self.device)
        for k, ar in enumerate((self.arena, self.arena_bwd)):
            for z in self._zreq[k]:
                z.ptr = ar.data_ptr() + 4 * z.offset
        self.loss5 = self.arena[self._loss5.offset:self._loss5.offset + 5]
        self.scalars = self.loss5[:4]

    # ---- allocation helpers ------------------------------------------------------------------------------
    def empty(self, shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def zeros(self, n, bwd=False):
        k = 1 if bwd else 0
        z = _Z(self._zsize[k], n)
        self._zsize[k] += -(-n // ALIGN) * ALIGN
        self._zreq[k].append(z)
        return z

    def set_precision(self, precision):
        """Switch every convolution descriptor between MVAE_PREC_FP32 and MVAE_PREC_TF32 (activations are shared)."""
        self.precision = precision
        for ops in self.enc_ops + self.dec_ops:
            for op in [b for o in ops for b in getattr(o, "blocks", [o])]:
                for name in ("desc", "d0", "d2"):
                    if hasattr(op, name):
                        getattr(op, name).precision = precision

    def new_T(self, shape, act):
        return T(self.empty(shape), self.empty(shape) if self.training else None, act)

    # ---- graph construction --------------------------------------------------------------------------------
    def _build(self):
        sp, B = self.spec, self.B
        L = sp.levels
        self.x = self.empty((B, sp.H, sp.W, sp.C))
        self.x_in = self.empty((B, sp.H, sp.W, sp.C)) if self.corrupt else self.x
        self.bands = [self.empty((B,) + sp.scales[i]) for i in range(L)]
        self.eps = [torch.zeros((B, z), dtype=torch.float32, device=self.device) for z in sp.z_dims]
        self.kl = self.empty((L, B))
        self.mulv, self.zT, self.ys = [], [], []
        self.enc_ops, self.dec_ops = [], []
        for i in range(L):
            ops = []
            p = f"encoder_{i}_"
            x = T(self.bands[i])
            op = Conv2D(self, x, p + "conv_base/kernel", p + "conv_base/bias", 3, 3, (1, 1), sp.conv_base_filters,
                        ACT_ELU, sp.coord, need_dx=False)
            ops.append(op)
            x, prev = op.y, sp.conv_base_filters
            for j, (f, k, s) in enumerate(sp.entries(sp.enc)):
                if s != (1, 1) or f != prev:
                    op = Conv2D(self, x, f"{p}_{j}_conv/kernel", f"{p}_{j}_conv/bias", k[0], k[1], s, f)
                    ops.append(op)
                    x = op.y
                op = MobileNetV3(self, x, f"{p}_{j}_mobilenetV3_", f)
                ops.append(op)
                x, prev = op.y, f
            h, w, c = sp.before_flatten[i]
            z = sp.z_dims[i]
            op = Conv2D(self, x.reshape(B, 1, 1, h * w * c), p + "mu_log_var/kernel", p + "mu_log_var/bias", 1, 1, (1, 1),
                        2 * z)
            ops.append(op)
            self.mulv.append(op.y)
            op = Reparam(self, op.y.reshape(B, 2 * z), self.eps[i], self.kl[i], z)
            ops.append(op)
            self.zT.append(op.y)
            self.enc_ops.append(fuse_chains(self, ops))
        for i in range(L):
            ops = []
            p = f"decoder_{i}_"
            h, w, c = sp.before_flatten[i]
            op = Conv2D(self, self.zT[i], p + "dense/kernel", p + "dense/bias", 1, 1, (1, 1), h * w * c,
                        need_dx=self.training)
            ops.append(op)
            x, prev = op.y.reshape(B, h, w, c), c
            for j, (f, k, s) in enumerate(sp.entries(sp.dec)):
                if s != (1, 1) or f != prev:
                    op = Conv2DTranspose(self, x, f"{p}_{j}_conv_transpose/kernel", f"{p}_{j}_conv_transpose/bias",
                                         k[0], k[1], s, f)
                    ops.append(op)
                    x = op.y
                op = MobileNetV3(self, x, f"{p}_{j}_mobilenetV3_", f)
                ops.append(op)
                x, prev = op.y, f
            op = Tail(self, x, p, sp.scales[i][2])
            ops.append(op)
            self.ys.append(op.y)
            self.dec_ops.append(fuse_chains(self, ops))
        lib = self.lib
        self.split_ws = self.empty((lib.mvae_pyramid_split_workspace_bytes(B, sp.H, sp.W, sp.C, L) // 4 + 1,))
        self.merge_ws = self.empty((lib.mvae_pyramid_merge_workspace_bytes(B, sp.H, sp.W, sp.C, L) // 4 + 1,))
        self.r0 = self.empty((B, sp.H, sp.W, sp.C))
        self.out = self.empty((B, sp.H, sp.W, sp.C))
        self.loss_sums = self.zeros(B * (1 + 2 * sp.C))
        self.per_sample = self.empty((3, B))
        # the five loss scalars of a step side by side in the forward arena (one device-to-host read): [0:4] written by
        # mvae_loss_finalize, [4] the regulariser sum mvae_optim_norms adds up
        self._loss5 = self.zeros(8)
        self.reg_loss = _Z(self._loss5.offset + 4, 1)
        self._zreq[0].append(self.reg_loss)
        self.sumsq = self.zeros(len(self.ps.segs))
        self.norm_partials = self.empty((2 * self.ps.nchunk,))      # one (|g|^2, regulariser) partial per optimiser chunk
        self.band_ptrs = (C.c_void_p * L)(*[b.data_ptr() for b in self.bands])
        self.y_ptrs = (C.c_void_p * L)(*[t.data.data_ptr() for t in self.ys])
        if self.training:
            self.dy_ptrs = (C.c_void_p * L)(*[t.grad.data_ptr() for t in self.ys])
        self.taps = (C.c_float * 9)(*[float(v) for v in sp.taps.ravel()])
        self.level_streams = None
        self._fork_wgrad, self._side_streams, self._side_used = False, {}, set()
        self.on_level_grads = None            # hook(level): set by the data-parallel wrapper
        self.current_level = 0
        self.on_dense_wgrad = None            # hook(op, level): a Dense weight gradient has just been issued (lane 0)
        self._deferred = {}
        self.defer_wgrad = os.environ.get("MVAE_NO_DEFER_WGRAD") != "1"
        self._flush_n = int(os.environ.get("MVAE_WGRAD_FLUSH_N", "1000"))
        self.defer_serial = False      # bench.py's per-call pass: same batched launches as the graph, on one stream
        # level-batched groups: position k of every level's op list (all levels are built from one config)
        # Opt-in (MVAE_BATCH_LEVELS=1): measured, one batched launch per layer is SLOWER than per-level launches on parallel
        # streams (cfg2 3.29 vs 3.09 ms, cfg3 10.1 vs 9.4 ms per step): the layers whose shape differs per level (conv_base,
        # Dense heads, tail) and the stride-2 dgrad still run per level and turn into join points of the single chain.
        self.batch_levels = os.environ.get("MVAE_BATCH_LEVELS") == "1"
        # Opt-in (MVAE_BATCH_COARSE=1): level 0 keeps its own chain of launches and the COARSE levels 1..L-1 run as ONE chain --
        # every step of their chains (fused tile kernels, gate kernels, strided convolutions, deferred weight gradients) is a
        # single multi-problem launch, and only the layers whose shape differs per level (conv_base, Dense, reparametrisation,
        # tail) fork to per-level streams.  Measured on cfg2: 293 -> 190 kernels per step and 4.5 -> 3.2 ms of summed kernel
        # time, but the single coarse chain (~130 dependent launches of ~10 us) ends later than four interleaved chains did
        # (1.46 vs 1.34 ms per step), so the default stays one chain per level.
        self.batch_coarse = os.environ.get("MVAE_BATCH_COARSE", "0") == "1" and sp.levels >= 3 and not self.batch_levels
        self._coarse = {}
        # first level of the single coarse chain (levels below it keep a chain of their own): MVAE_COARSE_FROM, default 1
        self.coarse_from = max(1, min(int(os.environ.get("MVAE_COARSE_FROM", "1")), sp.levels - 2))
        if self.batch_coarse:
            for name, lists in (("enc", self.enc_ops), ("dec", self.dec_ops)):
                if len({len(l) for l in lists}) != 1:
                    self.batch_coarse = False
                    break
                for k in range(len(lists[0])):
                    b = make_batched(self, [l[k] for l in lists[self.coarse_from:]])
                    if b is not None:
                        self._coarse[(name, k)] = b
        self._batched = {}
        for name, lists in (("enc", self.enc_ops), ("dec", self.dec_ops)):
            if len({len(l) for l in lists}) != 1:
                continue
            for k in range(len(lists[0])):
                b = make_batched(self, [l[k] for l in lists])
                if b is not None:
                    self._batched[(name, k)] = b

    # ---- execution ---------------------------------------------------------------------------------------------
    def _stream(self):
        self.s = torch.cuda.current_stream(self.device).cuda_stream

    def new_stream(self, priority=0):
        """A stream of this engine's own (never an alias of another branch's stream: _lib.OwnedStream); priority 0 lowest,
        1 middle, 2 / True highest.  The streams live as long as the process: captured graphs and pending work may refer to
        them after the engine is gone."""
        return _lib.OwnedStream.create(self.device, priority)

    def side(self, fn, lane=0, priority=0):
        """Weight-gradient launches: nothing later in the backward chain reads their output, so (in the multi-stream /
        CUDA-graph mode) they fork to a side stream of the current level and rejoin at the end of that level's backward;
        the dgrad chain -- the critical path -- never waits for them."""
        if not self._fork_wgrad:
            fn()
            return
        main = torch.cuda.current_stream(self.device)
        st = self._side_streams.get((main.cuda_stream, lane))
        if st is None:
            st = self._side_streams[(main.cuda_stream, lane)] = self.new_stream(priority)
        st.wait_stream(main)
        saved = self.s
        with torch.cuda.stream(st):
            self.s = st.cuda_stream
            fn()
        self.s = saved
        self._side_used.add((main.cuda_stream, lane))

    def wgrad(self, desc, x, gate, dy, dw, db):
        """A convolution weight gradient.  Nothing in the backward chain reads it, and every operand (saved activation,
        gradient buffer) stays valid until the step ends, so in the multi-stream / CUDA-graph mode the call is DEFERRED:
        the jobs of a level collect here and flush_wgrad() issues them as batched launches (same layer shape -> one launch
        through mvae_conv2d_wgrad_batched, CTAs split in proportion to the pixel counts).  Launched one by one next to the
        dgrad chain they cost ~0.5 ms of a 2.1 ms cfg2 step in SM contention; batched they stream at full width."""
        if os.environ.get("MVAE_DIAG_SKIP_WGRAD") == "1":      # diagnostic only (scripts/chain_probe.py): wrong gradients
            return
        # Dense heads (H = W = 1) are big single GEMMs at the START of a half of the backward: they go out at once and
        # overlap everything after them; only the convolutions' many small weight gradients are worth collecting
        if (self._fork_wgrad or self.defer_serial) and self.defer_wgrad and not (desc.H == 1 and desc.W == 1):
            key = torch.cuda.current_stream(self.device).cuda_stream
            self._deferred.setdefault(key, []).append((desc, x, gate, dy, dw, db))
            if len(self._deferred[key]) >= self._flush_n:
                self.flush_wgrad()
            return
        self.side(lambda: check(self.lib.mvae_conv2d_wgrad(C.byref(desc), x, gate, dy, dw, db, self.s), "conv2d_wgrad"))

    def flush_wgrad(self):
        """Issue the deferred weight gradients of the current stream, grouped by layer shape, on its side stream."""
        key = torch.cuda.current_stream(self.device).cuda_stream
        jobs = self._deferred.pop(key, [])
        if not jobs:
            return
        groups = {}
        for j in jobs:
            d = j[0]
            groups.setdefault((d.kh, d.kw, d.sh, d.sw, d.Cin, d.Cout, d.coord_mode, d.precision, d.H == 1 and d.W == 1),
                              []).append(j)

        def run(m):
            if len(m) == 1:
                d, x, gate, dy, dw, db = m[0]
                check(self.lib.mvae_conv2d_wgrad(C.byref(d), x, gate, dy, dw, db, self.s), "conv2d_wgrad")
            else:
                check(self.lib.mvae_conv2d_wgrad_batched(
                    len(m), _descs([j[0] for j in m]), _pa([j[1] for j in m]), _pa([j[2] for j in m]),
                    _pa([j[3] for j in m]), _pa([j[4] for j in m]), _pa([j[5] for j in m]), self.s),
                    "conv2d_wgrad_batched")

        # every launch on its own side stream: they are independent, and one alone rarely fills the GPU
        lane = 1
        for members in groups.values():
            for i in range(0, len(members), 8):
                self.side(lambda m=members[i:i + 8]: run(m), lane=lane)
                lane += 1

    def join_side(self):
        self.flush_wgrad()
        main = torch.cuda.current_stream(self.device)
        for key in [k for k in self._side_used if k[0] == main.cuda_stream]:
            main.wait_stream(self._side_streams[key])
            self._side_used.discard(key)

    def _levels(self, fn, parallel):
        """Run fn(i) for every level; with `parallel`, level i>0 goes to its own stream (fork/join, capturable)."""
        L = self.spec.levels
        if not parallel:
            self._stream()
            for i in range(L):
                fn(i)
            return
        self._ensure_streams()
        main = torch.cuda.current_stream(self.device)
        for i in range(L - 1, 0, -1):      # small levels first so they hide under level 0
            st = self.level_streams[i - 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                self._stream()
                fn(i)
        self.level0_stream.wait_stream(main)
        with torch.cuda.stream(self.level0_stream):
            self._stream()
            fn(0)
        main.wait_stream(self.level0_stream)
        for st in self.level_streams:
            main.wait_stream(st)
        self._stream()

    def _ensure_streams(self):
        if self.level_streams is None:
            L = self.spec.levels
            nhi = int(os.environ.get("MVAE_HIPRI_LEVELS", "1"))      # levels 0..nhi-1 on high-priority streams
            # stream priorities (measured on cfg2, ms / step): "three" 1.304 (default), "" = level 0 high, rest default 1.316,
            # "flat" 1.33, "coarse" = the coarse levels above level 0 1.42
            prio = os.environ.get("MVAE_PRIO", "three")
            if prio == "three":
                # three tiers: level 0 highest, the coarse levels in the middle, the weight-gradient side streams lowest
                # (their CTAs are scheduled only when nothing of a chain is waiting for an SM)
                self.level_streams = [self.new_stream(1) for i in range(L - 1)]
                nhi = 1
            elif prio == "coarse":
                self.level_streams = [self.new_stream(True) for i in range(L - 1)]
                nhi = 0
            elif prio == "flat":
                self.level_streams = [self.new_stream(False) for i in range(L - 1)]
                nhi = 0
            else:
                self.level_streams = [self.new_stream(i + 1 < nhi) for i in range(L - 1)]
            # level 0 is the critical path (75 % of the work, the longest chain): its kernels run on a high-priority
            # stream so their CTAs are never queued behind a coarse level's; the coarse levels fill the gaps
            self.level0_stream = self.new_stream(nhi >= 1)
            self.coarse_stream = self.new_stream()

    def _coarse_pass(self, method):
        """Levels coarse_from..L-1 as one chain on the current stream: position by position through their (identical) op
        lists; a position every level shares goes out as one multi-problem call, the rest fork to the level streams and
        rejoin."""
        L, c0 = self.spec.levels, self.coarse_from
        cur = torch.cuda.current_stream(self.device)
        fwd = method == "fwd"
        halves = (("enc", self.enc_ops), ("dec", self.dec_ops)) if fwd else (("dec", self.dec_ops), ("enc", self.enc_ops))
        for hi, (name, lists) in enumerate(halves):
            K = len(lists[0])
            for k in (range(K) if fwd else range(K - 1, -1, -1)):
                b = self._coarse.get((name, k))
                if b is not None:
                    self._stream()
                    getattr(b, method)()
                    continue
                for i in range(L - 1, c0 - 1, -1):
                    st = self.level_streams[i - 1]
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        self._stream()
                        getattr(lists[i][k], method)()
                        if not fwd:
                            self.join_side()          # a weight gradient this op deferred / forked rejoins its stream here
                for st in self.level_streams[c0 - 1:]:
                    cur.wait_stream(st)
                self._stream()
            if not fwd and hi == 0 and os.environ.get("MVAE_WGRAD_FLUSH_MID") == "1":
                self.flush_wgrad()                    # the decoders' weight gradients overlap the encoders' backward chain
        if not fwd:
            self.join_side()
            if self.on_level_grads is not None:
                for i in range(L - 1, c0 - 1, -1):
                    self.on_level_grads(i)

    def _use_coarse(self):
        # the multi-problem launches exist for the fused TF32 kernels; the fp32 mode keeps one chain per level
        return self.batch_coarse and self.precision == PREC_TF32 and self.fuse_mbv3 and bool(self._coarse)

    def _two_chains(self, fn, method):
        """Level 0 on its own (high-priority) stream, levels 1..coarse_from-1 on theirs (fn(i) runs a level), the coarse
        levels as one chain beside them; all rejoin the caller."""
        self._ensure_streams()
        main = torch.cuda.current_stream(self.device)
        self.coarse_stream.wait_stream(main)
        with torch.cuda.stream(self.coarse_stream):
            self._stream()
            self._coarse_pass(method)
        for i in range(self.coarse_from - 1, 0, -1):
            st = self.level_streams[i - 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                self._stream()
                fn(i)
        self.level0_stream.wait_stream(main)
        with torch.cuda.stream(self.level0_stream):
            self._stream()
            fn(0)
        main.wait_stream(self.level0_stream)
        for i in range(1, self.coarse_from):
            main.wait_stream(self.level_streams[i - 1])
        main.wait_stream(self.coarse_stream)
        self._stream()

    def _run_ops(self, name, lists, method, parallel):
        """Level-batched execution of one half (encoders / decoders): position by position; layers every level shares go
        out as one batched call on the current stream, the rest (conv_base, Dense heads, reparam, tail: shapes differ per
        level) fork to the level streams."""
        K = len(lists[0])
        order = range(K) if method == "fwd" else range(K - 1, -1, -1)
        for k in order:
            b = self._batched.get((name, k))
            if b is not None:
                self._stream()
                getattr(b, method)()
            else:
                def one(i, k=k):
                    getattr(lists[i][k], method)()
                    self.join_side()          # a weight gradient this op forked rejoins its level stream here

                self._levels(one, parallel)

    def split(self):
        sp = self.spec
        self._stream()
        check(self.lib.mvae_pyramid_split(_p(self.x_in), self.band_ptrs, _p(self.split_ws), self.B, sp.H, sp.W, sp.C,
                                          sp.levels, sp.v0, sp.v1, self.taps, 3, 3, sp.diff_mode, self.s), "pyramid_split")

    def corrupt_input(self, noise, keep, noise_std, rate):
        """multiscale_vae.py:139-147: x_in = GaussianNoise + SpatialDropout2D of the normalised x (noise ~ N(0,1) like x,
        keep (B,C) of 0/1; either may be None)."""
        sp = self.spec
        self._stream()
        check(self.lib.mvae_input_corrupt(_p(self.x), _p(noise), _p(keep), _p(self.x_in), self.B, sp.H * sp.W, sp.C, sp.v0,
                                          sp.v1, float(noise_std), 1.0 / (1.0 - rate) if keep is not None else 1.0, self.s),
              "input_corrupt")

    def zero_arena(self):
        self._stream()
        check(self.lib.mvae_memset_zero(self.arena.data_ptr(), self.arena.numel() * 4, self.s), "memset")

    def zero_bwd(self):
        self._stream()
        check(self.lib.mvae_memset_zero(self.arena_bwd.data_ptr(), self.arena_bwd.numel() * 4, self.s), "memset")
        check(self.lib.mvae_memset_zero(self.ps.grads.data_ptr(), self.ps.grads.numel() * 4, self.s), "memset")

    def forward_backward(self, parallel=False):
        """One training pass: gradients of mean_b(r*rf + kl*kf) land in ps.grads (regularisers are added by the
        optimiser kernel).  Inputs: self.x (raw image batch), self.eps[i]."""
        self.forward_train(parallel)
        self.backward(parallel)

    def forward_train(self, parallel=False):
        sp, lib, B = self.spec, self.lib, self.B
        self.zero_arena()
        self.split()

        def f(i):
            for op in self.enc_ops[i]:
                op.fwd()
            for op in self.dec_ops[i]:
                op.fwd()

        if parallel and self.batch_levels and self._batched:
            self._run_ops("enc", self.enc_ops, "fwd", parallel)
            self._run_ops("dec", self.dec_ops, "fwd", parallel)
            self._stream()
        elif parallel and self._use_coarse():
            self._two_chains(f, "fwd")
        else:
            self._levels(f, parallel)
        s = self.s
        check(lib.mvae_pyramid_merge_fwd(self.y_ptrs, _p(self.r0), _p(self.merge_ws), B, sp.H, sp.W, sp.C, sp.levels, s),
              "merge_fwd")
        check(lib.mvae_recon_loss_fwd(_p(self.r0), _p(self.x), _p(self.out), self.loss_sums.ptr, B, sp.H, sp.W, sp.C,
                                      sp.v0, sp.v1, s), "recon_loss_fwd")
        check(lib.mvae_loss_finalize(self.loss_sums.ptr, _p(self.kl), sp.levels, _p(self.per_sample), _p(self.scalars),
                                     B, sp.H, sp.W, sp.C, self.r_factor, self.kl_factor, s), "loss_finalize")

    def backward(self, parallel=False):
        sp, lib, B = self.spec, self.lib, self.B
        self.zero_bwd()
        s = self.s
        # d(loss)/d(r0) goes straight into the finest level's output gradient (dys[0] aliases dr0)
        check(lib.mvae_recon_loss_bwd(_p(self.r0), _p(self.x), self.loss_sums.ptr, _p(self.ys[0].grad), B, sp.H, sp.W,
                                      sp.C, sp.v0, sp.v1, self.r_factor / B, s), "recon_loss_bwd")
        check(lib.mvae_pyramid_merge_bwd(_p(self.ys[0].grad), self.dy_ptrs, B, sp.H, sp.W, sp.C, sp.levels, s), "merge_bwd")

        def g(i):
            self.current_level = i
            for op in reversed(self.dec_ops[i]):
                op.bwd()
            # The deferred weight gradients of the level go out when its chain is done (join_side below).  Issuing the decoder's
            # after the decoder half, next to the encoder's chain, measured slower on cfg2 (1.335 vs 1.316 ms / step): they
            # take SMs from the chain for work nothing waits for.  MVAE_WGRAD_FLUSH_MID=1 brings that back.
            if os.environ.get("MVAE_WGRAD_FLUSH_MID") == "1":
                self.flush_wgrad()
            for op in reversed(self.enc_ops[i]):
                op.bwd()
            # the level's chain is done: its last weight gradients are the tail of the step and take the whole GPU
            prev = self.lib.mvae_set_wgrad_sm_share(int(os.environ.get("MVAE_TAIL_WGRAD_SMS", "148")))
            try:
                self.join_side()
            finally:
                self.lib.mvae_set_wgrad_sm_share(prev)
            if self.on_level_grads is not None:
                self.on_level_grads(i)        # data parallel: this level's gradients are final -> exchange them now

        self._fork_wgrad = bool(parallel)
        try:
            if parallel and self.batch_levels and self._batched:
                self._run_ops("dec", self.dec_ops, "bwd", parallel)
                self._run_ops("enc", self.enc_ops, "bwd", parallel)
                self._stream()
                self.join_side()
            elif parallel and self._use_coarse():
                self._two_chains(g, "bwd")
            else:
                self._levels(g, parallel)
        finally:
            self._fork_wgrad = False

    def per_scale_elbo(self):
        """Per-scale ELBO terms of the LAST forward pass (self.x, self.ys, self.kl), in the form of
        multiscale_vae_.py:340-353: recon[i][b] = sum_hw mean_c |x_i - m_i| in raw units, with x_i the image at scale i
        (the low-pass chain of the split, multiscale_vae.py:292-315) and m_i the partial merge of the decoder outputs
        i..L-1 (multiscale_vae.py:210-219 stopped at scale i) denormalised and clipped (:221-222).  A metrics path, not
        part of the training step: every scale reuses the step's own kernels (split with i+1 levels ends in x_i, merge
        of the sub-pyramid gives m_i, the reconstruction-loss kernel sums |x_i - m_i|).  Returns (recon, kl), (L, B)."""
        sp, lib, B, L = self.spec, self.lib, self.B, self.spec.levels
        self._stream()
        s = self.s
        if not hasattr(self, "_pse"):
            self._pse = dict(bands=[self.empty((B,) + sp.scales[i]) for i in range(L)],
                             xraw=[self.empty((B,) + sp.scales[i]) for i in range(L)],
                             m=[self.empty((B,) + sp.scales[i]) for i in range(L)],
                             sums=torch.zeros((L, B, 1 + 2 * sp.C), dtype=torch.float32, device=self.device),
                             recon=self.empty((L, B)))
        P = self._pse
        check(lib.mvae_memset_zero(P["sums"].data_ptr(), P["sums"].numel() * 4, s), "memset")
        for i in range(L):
            h, w, c = sp.scales[i]
            # x_i: the last band of an (i+1)-level split is the low-passed, decimated image itself
            if i == 0:
                xraw = self.x                      # scale 0 is the raw image itself
            else:
                bp = (C.c_void_p * (i + 1))(*[b.data_ptr() for b in P["bands"][:i + 1]])
                check(lib.mvae_pyramid_split(_p(self.x), bp, _p(self.split_ws), B, sp.H, sp.W, sp.C, i + 1, sp.v0, sp.v1,
                                             self.taps, 3, 3, sp.diff_mode, s), "pyramid_split")
                check(lib.mvae_denormalize_clip(_p(P["bands"][i]), _p(P["xraw"][i]), P["bands"][i].numel(), sp.v0, sp.v1,
                                                s), "denormalize")
                xraw = P["xraw"][i]
            if i == L - 1:
                m = self.ys[i].data
            else:
                yp = (C.c_void_p * (L - i))(*[t.data.data_ptr() for t in self.ys[i:]])
                check(lib.mvae_pyramid_merge_fwd(yp, _p(P["m"][i]), _p(self.merge_ws), B, h, w, c, L - i, s), "merge_fwd")
                m = P["m"][i]
            check(lib.mvae_recon_loss_fwd(_p(m), _p(xraw), 0, P["sums"][i].data_ptr(), B, h, w, c, sp.v0, sp.v1, s),
                  "recon_loss_fwd")
        return P["sums"][:, :, 0] / float(sp.C), self.kl

    def optimizer_step(self, lr_dev, clip_norm, grad_scale=1.0):
        ps, lib = self.ps, self.lib
        self._stream()
        check(lib.mvae_optim_norms(ps.flat.data_ptr(), ps.grads.data_ptr(), ps.seg_table.data_ptr(),
                                   ps.chunk_table.data_ptr(), ps.nseg, ps.nchunk, CHUNK, grad_scale,
                                   _p(self.norm_partials), self.sumsq.ptr, self.reg_loss.ptr, self.s), "optim_norms")
        check(lib.mvae_optim_adagrad(ps.flat.data_ptr(), ps.grads.data_ptr(), ps.acc.data_ptr(), ps.seg_table.data_ptr(),
                                     ps.chunk_table.data_ptr(), ps.nchunk, CHUNK, self.sumsq.ptr, lr_dev.data_ptr(),
                                     float(clip_norm) if clip_norm else 0.0, 1e-7, self.s), "optim_adagrad")

    # ---- inference paths -----------------------------------------------------------------------------------
    def encode(self, parallel=False):
        """`_model_encoder` (multiscale_vae.py:228-243): x -> pyramid -> per-level sampled z."""
        self.zero_arena()
        self.split()

        def f(i):
            for op in self.enc_ops[i]:
                op.fwd()

        self._levels(f, parallel)

    def decode(self, parallel=False):
        """`_model_decoder` (multiscale_vae.py:247-257): self.zT[i].data -> merged, denormalised image in self.out."""
        sp, lib, B = self.spec, self.lib, self.B
        self.zero_arena()

        def f(i):
            for op in self.dec_ops[i]:
                op.fwd()

        self._levels(f, parallel)
        s = self.s
        check(lib.mvae_pyramid_merge_fwd(self.y_ptrs, _p(self.r0), _p(self.merge_ws), B, sp.H, sp.W, sp.C, sp.levels, s),
              "merge_fwd")
        check(lib.mvae_denormalize_clip(_p(self.r0), _p(self.out), self.r0.numel(), sp.v0, sp.v1, s), "denormalize")
