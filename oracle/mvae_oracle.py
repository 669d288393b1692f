"""CPU oracle for the multiscale-VAE training step.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU with plain PyTorch ops (fp32 or fp64), the graph that
`/root/reference/mvae/multiscale_vae.py` builds out of Keras layers.  It is the parity checker
for the CUDA path and the CPU baseline timed by `bench.py`; nothing in the product package
(`multiscale_variational_autoencoder_b200/`) may import it.

PARITY STATUS
  * pinned by the reference's own tests / code executed here:
      - `gaussian_kernel`            vs the reference function executed through `oracle/ref_shim.py`
                                     (golden taps in tests/golden/gaussian_kernel.npz)
      - Gaussian filter zeros/ones   `tests/test_layer_blocks.py:9-39`
      - Laplacian split->merge       `tests/test_layer_blocks.py:160-190`
      - CoordinateChannel2D values   reference `coord.py:88-133` executed on a numpy Keras-backend shim
  * PARITY UNPINNED (TensorFlow 2.3.1 / Keras 2.4.3 are not installable here, the reference holds no
    golden vectors): conv / conv-transpose SAME alignment, bilinear resize convention, BatchNorm,
    hard_sigmoid, Adagrad/clipnorm, regulariser constants.  Those follow the cited source lines plus
    documented TF semantics.

Layouts follow Keras: activations NHWC, Conv2D kernel (kh,kw,Cin,Cout), Conv2DTranspose kernel
(kh,kw,Cout,Cin), depthwise kernel (kh,kw,C,1), Dense kernel (in,out), Flatten order (h,w,c).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# Training-phase input corruption                                     (multiscale_vae.py:58-59, 139-147)
# ---------------------------------------------------------------------------------------------


def corrupt_normalized(x, noise, keep, v0, v1, noise_std, rate):
    """GaussianNoise(stddev=noise_std) then SpatialDropout2D(rate) applied to the NORMALISED image, as the "multiscale"
    model does in the training phase (multiscale_vae.py:136-147).  noise ~ N(0,1) shaped like x and keep (B,C) of 0/1 are
    supplied so that the result is deterministic; returns the normalised, corrupted image (B,H,W,C)."""
    t = 2.0 * (x - v0) / (v1 - v0) - 1.0                       # normalize Lambda, :79-84
    if noise is not None:
        t = t + noise_std * noise                               # keras.layers.GaussianNoise, :141-142
    if keep is not None:
        t = t * keep[:, None, None, :] / (1.0 - rate)            # keras.layers.SpatialDropout2D, :146-147
    return t


# ---------------------------------------------------------------------------------------------
# Gaussian kernel / filter                                  (layer_blocks.py:980-1002, 1008-1050)
# ---------------------------------------------------------------------------------------------


def gaussian_kernel(size, nsig):
    """2-D Gaussian taps, fp64.  Restates layer_blocks.py:980-1002."""
    assert len(size) == 2 and len(nsig) == 2
    axes = [np.linspace(-abs(nsig[i]), abs(nsig[i]), size[i], endpoint=True) for i in range(2)]
    gx, gy = np.meshgrid(axes[0], axes[1])
    d = np.sqrt(gx * gx + gy * gy)
    g = np.exp(-(d ** 2) / 2.0)
    return g / g.sum()


def _taps(kernel_size, nsig, dtype):
    # Keras stores the initialiser output in a float32 variable (layer_blocks.py:1029-1037)
    k32 = gaussian_kernel(kernel_size, nsig).astype(np.float32)
    return torch.from_numpy(k32).to(dtype)


def gaussian_filter(x, kernel_size=(3, 3), nsig=(1, 1)):
    """Frozen depthwise conv, SAME zero padding, stride 1, no bias (layer_blocks.py:1039-1050).

    x: (B,H,W,C) tensor.  The same taps are applied to every channel (layer_blocks.py:1035-1036)."""
    B, H, W, C = x.shape
    k = _taps(kernel_size, nsig, x.dtype)
    kh, kw = k.shape
    w = k.reshape(1, 1, kh, kw).repeat(C, 1, 1, 1)
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, ((kw - 1) // 2, kw // 2, (kh - 1) // 2, kh // 2))
    y = F.conv2d(xn, w, groups=C)
    return y.permute(0, 2, 3, 1)


# ---------------------------------------------------------------------------------------------
# normalise / denormalise / bilinear / pyramid      (multiscale_vae.py:79-94,129-160,204-224,292-315)
# ---------------------------------------------------------------------------------------------


def normalize(y, v0, v1):
    """multiscale_vae.py:79-84"""
    return 2.0 * (y - v0) / (v1 - v0) - 1.0


def denormalize(y, v0, v1):
    """multiscale_vae.py:86-94"""
    return torch.clamp((y + 1.0) * (v1 - v0) / 2.0 + v0, min=v0, max=v1)


def bilinear_up2(x):
    """UpSampling2D(size=2, interpolation='bilinear') == tf.image.resize half-pixel centres, edge clamp
    (multiscale_vae.py:214-216)."""
    y = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False)
    return y.permute(0, 2, 3, 1)


def decimate2(x):
    """MaxPool2D(pool 1x1, stride 2, VALID) == x[:, ::2, ::2, :] (multiscale_vae.py:308-311)."""
    return x[:, ::2, ::2, :]


def pyramid_split(x, levels, v0=0.0, v1=255.0, nsig=(2, 2), kernel_size=(3, 3), mode="no_upsample"):
    """Band-pass pyramid of a raw image.

    mode "no_upsample": multiscale_vae.py:129-160 + 292-315 (diff = i0 - gauss(i0)).
    mode "laplacian" : layer_blocks.py:23-101 (diff = i0 - up2(down2(gauss(i0))))."""
    layer = normalize(x, v0, v1)
    out = []
    for i in range(levels):
        if i == levels - 1:
            out.append(layer)
        else:
            f0 = gaussian_filter(layer, kernel_size, nsig)
            d0 = decimate2(f0)
            if mode == "no_upsample":
                diff = layer - f0
            elif mode == "laplacian":
                diff = layer - bilinear_up2(d0)
            else:
                raise ValueError(mode)
            out.append(diff)
            layer = d0
    return out


def pyramid_scales(x, levels, v0=0.0, v1=255.0, nsig=(2, 2), kernel_size=(3, 3)):
    """The low-pass chain of the split (multiscale_vae.py:292-315: f0 = gauss(i0), d0 = f0[::2, ::2]): normalised image at
    every scale, x_0 = normalize(x), x_{i+1} = decimate(gauss(x_i)).  x_{L-1} is the last output of pyramid_split."""
    layer = normalize(x, v0, v1)
    out = [layer]
    for _ in range(levels - 1):
        layer = decimate2(gaussian_filter(layer, kernel_size, nsig))
        out.append(layer)
    return out


def pyramid_merge_raw(ys):
    """Coarse-to-fine bilinear x2 + add (multiscale_vae.py:204-219), before denormalisation."""
    r = ys[-1]
    for i in range(len(ys) - 2, -1, -1):
        r = bilinear_up2(r) + ys[i]
    return r


def pyramid_merge(ys, v0=0.0, v1=255.0):
    """multiscale_vae.py:204-224 / layer_blocks.py:107-185 (trainable=False branch)."""
    return denormalize(pyramid_merge_raw(ys), v0, v1)


# ---------------------------------------------------------------------------------------------
# CoordConv channels                                                        (coord.py:88-133)
# ---------------------------------------------------------------------------------------------


def coordinate_channels_2d(x, use_radius=False):
    """Append xx (row index), yy (column index) in [-1,1] and optionally rr (coord.py:88-133)."""
    B, H, W, C = x.shape
    ii = torch.arange(H, dtype=x.dtype).view(1, H, 1, 1).expand(B, H, W, 1)
    jj = torch.arange(W, dtype=x.dtype).view(1, 1, W, 1).expand(B, H, W, 1)
    xx = ii / (H - 1) * 2 - 1.0          # coord.py:117-119
    yy = jj / (W - 1) * 2 - 1.0          # coord.py:121-123
    out = [x, xx, yy]
    if use_radius:
        out.append(torch.sqrt((xx - 0.5) ** 2 + (yy - 0.5) ** 2))   # coord.py:127-130
    return torch.cat(out, dim=-1)


# ---------------------------------------------------------------------------------------------
# Conv helpers with TensorFlow SAME semantics                               
# ---------------------------------------------------------------------------------------------


def same_pads(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2, out


def conv2d_same(x, w, b, stride=(1, 1)):
    """Keras Conv2D(padding='same').  w: (kh,kw,Cin,Cout)."""
    kh, kw = w.shape[0], w.shape[1]
    pt, pb, _ = same_pads(x.shape[1], kh, stride[0])
    pl, pr, _ = same_pads(x.shape[2], kw, stride[1])
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, w, b, stride=(1, 1)):
    """Keras Conv2DTranspose(padding='same'): the gradient of the SAME forward conv that maps
    (H*s, W*s) -> (H, W).  w: (kh,kw,Cout,Cin)."""
    kh, kw = w.shape[0], w.shape[1]
    H, W = x.shape[1], x.shape[2]
    pt, _, _ = same_pads(H * stride[0], kh, stride[0])
    pl, _, _ = same_pads(W * stride[1], kw, stride[1])
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), None, stride=stride)
    # F.pad so that the crop below never runs out of rows/cols (k < s cases)
    need_h, need_w = pt + H * stride[0], pl + W * stride[1]
    y = F.pad(y, (0, max(need_w - y.shape[3], 0), 0, max(need_h - y.shape[2], 0)))
    y = y[:, :, pt:pt + H * stride[0], pl:pl + W * stride[1]]
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1)


def depthwise_same(x, w, b):
    """Keras DepthwiseConv2D 3x3 stride 1 SAME, depth_multiplier 1.  w: (kh,kw,C,1)."""
    C = x.shape[3]
    kh, kw = w.shape[0], w.shape[1]
    xn = F.pad(x.permute(0, 3, 1, 2), ((kw - 1) // 2, kw // 2, (kh - 1) // 2, kh // 2))
    y = F.conv2d(xn, w.permute(2, 3, 0, 1), b, groups=C)
    return y.permute(0, 2, 3, 1)


def hard_sigmoid(x):
    """Keras <= 2.x hard_sigmoid = clip(0.2 x + 0.5, 0, 1)."""
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def batchnorm_train(x, gamma, beta, eps, axes):
    """Training-mode BatchNormalization: biased batch variance."""
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    return gamma * (x - mean) / torch.sqrt(var + eps) + beta, mean.flatten(), var.flatten()


def batchnorm_infer(x, gamma, beta, mean, var, eps):
    return gamma * (x - mean) / torch.sqrt(var + eps) + beta


# ---------------------------------------------------------------------------------------------
# Model description shared by oracle construction
# ---------------------------------------------------------------------------------------------

REG_NONE, REG_L1, REG_L2 = 0, 1, 2
REG_FACTOR = 0.01       # Keras string regularisers 'l1' / 'l2' use factor 0.01
SE_BN_EPS, SE_BN_MOM = 1e-3, 0.99          # Keras BatchNormalization defaults (layer_blocks.py:447-449)
DEC_BN_EPS, DEC_BN_MOM = 1e-4, 0.999       # multiscale_vae.py:420-421


def glorot_normal_(t, fan_in, fan_out, gen):
    """Keras glorot_normal: truncated normal (+-2 sigma), sigma = sqrt(2/(fan_in+fan_out))/0.8796..."""
    std = math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return t


class OracleMVAE:
    """Restatement of `MultiscaleVAE` (multiscale_vae.py:11-587) as explicit tensor code.

    Parameters live in `self.params` (OrderedDict name -> tensor), `self.reg` (name -> REG_*),
    `self.trainable` (name -> bool).  Non-trainable entries are the BatchNorm moving statistics."""

    def __init__(self, input_dims, z_dims, encoder, decoder=None, min_value=0.0, max_value=255.0,
                 sample_std=0.01, channels_index=2, coord_conv=None, logvar_scale=1.0,
                 diff_mode="no_upsample", dtype=torch.float32, seed=7):
        if encoder is None:
            raise ValueError("encoder cannot be None")                 # multiscale_vae.py:35-36
        if not all(i > 0 for i in z_dims):
            raise ValueError("z_dims elements should be > 0")          # multiscale_vae.py:37-38
        if decoder is None:                                            # multiscale_vae.py:40-45
            decoder = {k: encoder[k][::-1] for k in ("filters", "strides", "kernel_size")}
        assert channels_index == 2
        self.input_dims = tuple(input_dims)
        self.z_dims = list(z_dims)
        self.levels = len(z_dims)
        self.enc_cfg, self.dec_cfg = encoder, decoder
        self.v0, self.v1 = float(min_value), float(max_value)
        self.sample_std = float(sample_std)
        self.coord_conv = coord_conv
        self.logvar_scale = float(logvar_scale)
        self.diff_mode = diff_mode
        self.dtype = dtype
        self.conv_base_filters = 32                                     # multiscale_vae.py:50
        self.nsig, self.gk = (2, 2), (3, 3)                             # multiscale_vae.py:56-57
        H, W, C = self.input_dims
        self.scales = [(H, W, C)]
        for _ in range(1, self.levels):                                 # multiscale_vae.py:111-127
            h, w, c = self.scales[-1]
            self.scales.append((int(h / 2), int(w / 2), c))
        self.params, self.reg, self.trainable = OrderedDict(), {}, {}
        self.shape_before_flatten = []
        gen = torch.Generator().manual_seed(seed)
        self._gen = gen
        for i in range(self.levels):
            self._build_encoder(i)
        for i in range(self.levels):
            self._build_decoder(i)
        self.acc = None
        self.lr, self.r_factor, self.kl_factor, self.clip_norm = 0.01, 1.0, 1.0, 1.0

    # ---- parameter creation ------------------------------------------------------------------
    def _add(self, name, shape, reg=REG_NONE, fans=None, value=None, trainable=True):
        t = torch.zeros(shape, dtype=self.dtype)
        if fans is not None:
            glorot_normal_(t, fans[0], fans[1], self._gen)
        if value is not None:
            t.fill_(value)
        self.params[name] = t
        self.reg[name] = reg
        self.trainable[name] = trainable
        return t

    def _add_conv(self, name, kh, kw, cin, cout, reg, transpose=False):
        shape = (kh, kw, cout, cin) if transpose else (kh, kw, cin, cout)
        # Keras fans for a 4-D kernel: receptive * shape[-2], receptive * shape[-1]
        self._add(name + "/kernel", shape, reg, fans=(kh * kw * shape[2], kh * kw * shape[3]))
        self._add(name + "/bias", (cout,))

    def _add_dense(self, name, kin, kout, reg):
        self._add(name + "/kernel", (kin, kout), reg, fans=(kin, kout))
        self._add(name + "/bias", (kout,))

    def _add_bn(self, name, c):
        self._add(name + "/gamma", (c,), value=1.0)
        self._add(name + "/beta", (c,))
        self._add(name + "/moving_mean", (c,), trainable=False)
        self._add(name + "/moving_variance", (c,), value=1.0, trainable=False)

    def _add_mbv3(self, prefix, cin, filters):
        """layer_blocks.py:556-648 (+ squeeze_excite_block 418-462 with use_batchnorm=True)."""
        self._add_conv(prefix + "conv0", 1, 1, cin, filters, REG_L1)
        self._add(prefix + "conv1/depthwise_kernel", (3, 3, filters, 1), REG_L1,
                  fans=(3 * 3 * filters, 3 * 3 * 1))
        self._add(prefix + "conv1/bias", (filters,))
        se = prefix + "squeeze_excite_"
        self._add_dense(se + "dense0", filters, filters, REG_L1)
        self._add_bn(se + "batchnorm0", filters)
        self._add_dense(se + "dense1", filters, filters, REG_L1)
        self._add_conv(prefix + "conv2", 1, 1, filters, cin, REG_L1)

    def _block_entries(self, cfg):
        f, k, s = cfg["filters"], cfg["kernel_size"], cfg["strides"]
        if len(f) != len(k) or len(f) != len(s) or len(f) <= 0:       # layer_blocks.py:918-924
            raise ValueError("len(filters) should be equal to len(kernel_size) and len(strides)")
        return list(zip(f, k, s))

    def _build_encoder(self, i):
        """multiscale_vae.py:319-385"""
        h, w, c = self.scales[i]
        p = f"encoder_{i}_"
        cin = c + {None: 0, "xy": 2, "xyr": 3}[self.coord_conv]
        self._add_conv(p + "conv_base", 3, 3, cin, self.conv_base_filters, REG_L2)
        prev = self.conv_base_filters
        for j, (f, k, s) in enumerate(self._block_entries(self.enc_cfg)):
            if s[0] != 1 or s[1] != 1 or f != prev:                      # layer_blocks.py:946-949
                self._add_conv(f"{p}_{j}_conv", k[0], k[1], prev, f, REG_L1)
                h, w = -(-h // s[0]), -(-w // s[1])
            self._add_mbv3(f"{p}_{j}_mobilenetV3_", f, f)
            prev = f
        self.shape_before_flatten.append((h, w, prev))
        K = h * w * prev
        self._add_dense(p + "mu", K, self.z_dims[i], REG_L2)
        self._add_dense(p + "log_var", K, self.z_dims[i], REG_L2)

    def _build_decoder(self, i):
        """multiscale_vae.py:389-433"""
        h, w, c = self.shape_before_flatten[i]
        p = f"decoder_{i}_"
        self._add_dense(p + "dense", self.z_dims[i], h * w * c, REG_L2)
        prev = c
        for j, (f, k, s) in enumerate(self._block_entries(self.dec_cfg)):
            if s[0] != 1 or s[1] != 1 or f != prev:                      # layer_blocks.py:950-951
                self._add_conv(f"{p}_{j}_conv_transpose", k[0], k[1], prev, f, REG_L1, transpose=True)
                h, w = h * s[0], w * s[1]
            self._add_mbv3(f"{p}_{j}_mobilenetV3_", f, f)
            prev = f
        if (h, w) != self.scales[i][:2]:
            raise ValueError(f"level {i}: decoder emits {h}x{w}, scale is {self.scales[i][:2]} "
                             "(total encoder stride must divide every scale; SURVEY App. C-7)")
        self._add_bn(p + "batchnorm", prev)
        self._add_conv(p + "conv_out", 1, 1, prev, self.scales[i][2], REG_L2)

    # ---- forward pieces ----------------------------------------------------------------------
    def _se(self, u, prefix, training, new_stats):
        """layer_blocks.py:418-462"""
        P = self.params
        g = u.mean(dim=(1, 2))
        g = self._relu(g @ P[prefix + "dense0/kernel"] + P[prefix + "dense0/bias"], prefix + "h")
        bn = prefix + "batchnorm0"
        if training:
            g, m, v = batchnorm_train(g, P[bn + "/gamma"], P[bn + "/beta"], SE_BN_EPS, (0,))
            new_stats[bn] = (m.detach(), v.detach(), SE_BN_MOM, 1.0)   # 2-D input: biased variance
        else:
            g = batchnorm_infer(g, P[bn + "/gamma"], P[bn + "/beta"], P[bn + "/moving_mean"],
                                P[bn + "/moving_variance"], SE_BN_EPS)
        g = g @ P[prefix + "dense1/kernel"] + P[prefix + "dense1/bias"]
        m = None if self.masks is None else self.masks.get(prefix + "hs")
        if m is None:
            g = hard_sigmoid(g)
        else:       # the product's pass mask: linear where it passed, the saturated constant elsewhere
            g = torch.where(m, 0.2 * g + 0.5, hard_sigmoid(g).detach())
        return g[:, None, None, :] * u

    masks = None

    def _relu(self, pre, key):
        """relu, or -- when `self.masks` carries the product's activation pattern for this tensor -- the same linear piece
        the product took (tests: gradients of a TF32 run are only comparable on the SAME piecewise-linear network; a
        pre-activation within TF32 rounding of the kink otherwise flips an O(1) gradient entry)."""
        m = None if self.masks is None else self.masks.get(key)
        return torch.relu(pre) if m is None else torch.where(m, pre, torch.zeros_like(pre))

    def _mbv3(self, x, prefix, training, new_stats):
        """layer_blocks.py:594-641"""
        P = self.params
        a = self._relu(conv2d_same(x, P[prefix + "conv0/kernel"], P[prefix + "conv0/bias"]), prefix + "a")
        u = self._relu(depthwise_same(a, P[prefix + "conv1/depthwise_kernel"], P[prefix + "conv1/bias"]), prefix + "u")
        v = self._se(u, prefix + "squeeze_excite_", training, new_stats)
        y = conv2d_same(v, P[prefix + "conv2/kernel"], P[prefix + "conv2/bias"])
        return y + x

    def encode_level(self, i, band, eps, training=True, new_stats=None, taps=None):
        """multiscale_vae.py:319-385.  eps ~ N(0,1); the reference draws N(0, sample_std^2)."""
        P = self.params
        new_stats = {} if new_stats is None else new_stats
        p = f"encoder_{i}_"
        x = band
        if self.coord_conv is not None:
            x = coordinate_channels_2d(x, use_radius=(self.coord_conv == "xyr"))
        x = F.elu(conv2d_same(x, P[p + "conv_base/kernel"], P[p + "conv_base/bias"]))
        if taps is not None:
            taps[p + "conv_base"] = x
        prev = self.conv_base_filters
        for j, (f, k, s) in enumerate(self._block_entries(self.enc_cfg)):
            if s[0] != 1 or s[1] != 1 or f != prev:
                x = conv2d_same(x, P[f"{p}_{j}_conv/kernel"], P[f"{p}_{j}_conv/bias"], stride=s)
                if taps is not None:
                    taps[f"{p}_{j}_conv"] = x
            x = self._mbv3(x, f"{p}_{j}_mobilenetV3_", training, new_stats)
            if taps is not None:
                taps[f"{p}_{j}_mobilenetV3_add"] = x
            prev = f
        flat = x.reshape(x.shape[0], -1)
        mu = flat @ P[p + "mu/kernel"] + P[p + "mu/bias"]
        log_var = flat @ P[p + "log_var/kernel"] + P[p + "log_var/bias"]
        # multiscale_vae.py:372-378 (logvar_scale 1.0) / multiscale_vae_.py:34 (0.5)
        z = mu + torch.exp(self.logvar_scale * log_var) * (self.sample_std * eps)
        return z, mu, log_var

    def decode_level(self, i, z, training=True, new_stats=None, taps=None):
        """multiscale_vae.py:389-433"""
        P = self.params
        new_stats = {} if new_stats is None else new_stats
        p = f"decoder_{i}_"
        h, w, c = self.shape_before_flatten[i]
        x = (z @ P[p + "dense/kernel"] + P[p + "dense/bias"]).reshape(-1, h, w, c)
        prev = c
        for j, (f, k, s) in enumerate(self._block_entries(self.dec_cfg)):
            if s[0] != 1 or s[1] != 1 or f != prev:
                x = conv2d_transpose_same(x, P[f"{p}_{j}_conv_transpose/kernel"],
                                          P[f"{p}_{j}_conv_transpose/bias"], stride=s)
                if taps is not None:
                    taps[f"{p}_{j}_conv_transpose"] = x
            x = self._mbv3(x, f"{p}_{j}_mobilenetV3_", training, new_stats)
            if taps is not None:
                taps[f"{p}_{j}_mobilenetV3_add"] = x
            prev = f
        bn = p + "batchnorm"
        if training:
            x, m, v = batchnorm_train(x, P[bn + "/gamma"], P[bn + "/beta"], DEC_BN_EPS, (0, 1, 2))
            n = x.shape[0] * x.shape[1] * x.shape[2]
            # fused 4-D BatchNorm feeds the unbiased variance into the moving average
            new_stats[bn] = (m.detach(), v.detach(), DEC_BN_MOM, n / max(n - 1, 1))
        else:
            x = batchnorm_infer(x, P[bn + "/gamma"], P[bn + "/beta"], P[bn + "/moving_mean"],
                                P[bn + "/moving_variance"], DEC_BN_EPS)
        return conv2d_same(x, P[p + "conv_out/kernel"], P[p + "conv_out/bias"])

    def split_bands(self, x):
        return pyramid_split(x, self.levels, self.v0, self.v1, self.nsig, self.gk, self.diff_mode)

    # ---- losses (multiscale_vae.py:453-495) ---------------------------------------------------
    def r_loss_metric(self, y, yh):
        return (y - yh).abs().mean(dim=(1, 2, 3))                       # :453-456

    def r_loss(self, y, yh):
        H, W, _ = self.input_dims
        d0, d1 = int(H / 2), int(W / 2)
        px = (y - yh).abs().mean(dim=(1, 2, 3))
        ch = (y.mean(dim=(1, 2)) - yh.mean(dim=(1, 2))).abs()
        r0, r1, c0, c1 = int(d0 / 2), int(d0 * 3 / 2), int(d1 / 2), int(d1 * 3 / 2)
        cc = (y[:, r0:r1, c0:c1, :].mean(dim=(1, 2)) - yh[:, r0:r1, c0:c1, :].mean(dim=(1, 2))).abs()
        return px + (ch.mean(dim=1) + cc.mean(dim=1)) / 2.0             # :458-481

    @staticmethod
    def kl_loss(mu, log_var):
        return -0.5 * (1.0 + log_var - mu ** 2 - torch.exp(log_var)).sum(dim=-1)   # :485-488

    def reg_loss(self):
        tot = 0.0
        for n, t in self.params.items():
            if self.reg[n] == REG_L1:
                tot = tot + REG_FACTOR * t.abs().sum()
            elif self.reg[n] == REG_L2:
                tot = tot + REG_FACTOR * (t ** 2).sum()
        return tot

    # ---- whole graph -------------------------------------------------------------------------
    def forward(self, x, eps, training=True, taps=None):
        """`_model_trainable` (multiscale_vae.py:261-288) with training noise/dropout disabled.

        x: (B,H,W,C) raw image; eps: list of (B,z_i) standard-normal tensors.  Returns a dict."""
        new_stats = {}
        bands = self.split_bands(x)
        zs, mus, lvs, ys = [], [], [], []
        for i in range(self.levels):
            z, mu, lv = self.encode_level(i, bands[i], eps[i], training, new_stats, taps)
            zs.append(z), mus.append(mu), lvs.append(lv)
            ys.append(self.decode_level(i, z, training, new_stats, taps))
        out = pyramid_merge(ys, self.v0, self.v1)
        if self.masks is not None and "out_clip" in self.masks:
            # the product's clip pattern of the denormalize Lambda and the sign pattern of |y - y_hat|
            raw = (pyramid_merge_raw(ys) + 1.0) * (self.v1 - self.v0) / 2.0 + self.v0
            out = torch.where(self.masks["out_clip"], raw, out.detach())
        mu, lv = torch.cat(mus, -1), torch.cat(lvs, -1)
        r = self.r_loss(x, out)
        if self.masks is not None and "l1_sign" in self.masks:
            d0, d1 = int(self.input_dims[0] / 2), int(self.input_dims[1] / 2)
            px = ((x - out) * self.masks["l1_sign"]).mean(dim=(1, 2, 3))
            r = r - (x - out).abs().mean(dim=(1, 2, 3)) + px
        kl = self.kl_loss(mu, lv)
        reg = self.reg_loss()
        loss = (r * self.r_factor + kl * self.kl_factor).mean() + reg
        return dict(bands=bands, z=zs, mu=mus, log_var=lvs, y=ys, out=out, r_loss=r, kl_loss=kl,
                    kl_per_scale=[self.kl_loss(m, l) for m, l in zip(mus, lvs)],
                    r_metric=self.r_loss_metric(x, out), reg_loss=reg, loss=loss, new_stats=new_stats)

    def per_scale_elbo(self, x, eps, training=True):
        """Per-scale ELBO terms in the form of multiscale_vae_.py:345-353 (the north star's "per-scale ELBO"):
            recon_i = mean_b sum_hw MAE_c( data_scale_i, reconstruction_i ),   kl_i = mean_b KL_i   (:340-342)
        on THIS model's pyramid: data_scale_i = the image at scale i (the split's low-pass chain, raw units),
        reconstruction_i = the partial merge of the decoder outputs i..L-1 (multiscale_vae.py:210-219 stopped at scale
        i), denormalised and clipped (:221-222).  Returns per-sample tensors (L, B)."""
        res = self.forward(x, eps, training=training)
        xs = pyramid_scales(x, self.levels, self.v0, self.v1, self.nsig, self.gk)
        recon = []
        for i in range(self.levels):
            target = denormalize(xs[i], self.v0, self.v1)
            rec = pyramid_merge(res["y"][i:], self.v0, self.v1)
            recon.append((target - rec).abs().mean(dim=3).sum(dim=(1, 2)))
        recon, kl = torch.stack(recon), torch.stack(res["kl_per_scale"])
        return dict(recon=recon, kl=kl, elbo=recon * self.r_factor + kl * self.kl_factor, res=res)

    def encode(self, x, eps):
        """`_model_encoder` (multiscale_vae.py:228-243): sampled z of every level, concatenated."""
        bands = self.split_bands(x)
        return torch.cat([self.encode_level(i, bands[i], eps[i], training=False)[0]
                          for i in range(self.levels)], -1)

    def decode(self, z):
        """`_model_decoder` (multiscale_vae.py:247-257)."""
        zs = torch.split(z, self.z_dims, dim=-1)
        return pyramid_merge([self.decode_level(i, zs[i], training=False) for i in range(self.levels)],
                             self.v0, self.v1)

    # ---- training step (compile + fit; multiscale_vae.py:437-504, 550) --------------------------
    def compile(self, learning_rate, r_loss_factor=1.0, kl_loss_factor=1.0, clip_norm=1.0):
        self.lr, self.r_factor, self.kl_factor, self.clip_norm = \
            learning_rate, r_loss_factor, kl_loss_factor, clip_norm
        self.acc = {n: torch.full_like(t, 0.1) for n, t in self.params.items() if self.trainable[n]}

    def loss_and_grads(self, x, eps):
        names = [n for n in self.params if self.trainable[n]]
        for n in names:
            self.params[n].requires_grad_(True)
            self.params[n].grad = None
        res = self.forward(x, eps, training=True)
        res["loss"].backward()
        grads = {n: (self.params[n].grad if self.params[n].grad is not None
                     else torch.zeros_like(self.params[n])) for n in names}
        for n in names:
            self.params[n].requires_grad_(False)
        return res, grads

    def apply_grads(self, grads, new_stats):
        """Keras Adagrad(initial_accumulator 0.1, eps 1e-7) with per-variable clipnorm (App. A.13)."""
        with torch.no_grad():
            for n, g in grads.items():
                if self.clip_norm is not None:
                    nrm = g.norm()
                    g = g * (self.clip_norm / torch.clamp(nrm, min=self.clip_norm))
                self.acc[n] += g * g
                self.params[n] -= self.lr * g / (torch.sqrt(self.acc[n]) + 1e-7)
            for bn, (m, v, mom, corr) in new_stats.items():
                self.params[bn + "/moving_mean"].mul_(mom).add_(m * (1 - mom))
                self.params[bn + "/moving_variance"].mul_(mom).add_(v * corr * (1 - mom))

    def train_step(self, x, eps):
        res, grads = self.loss_and_grads(x, eps)
        self.apply_grads(grads, res["new_stats"])
        return res, grads

    # ---- state dict ---------------------------------------------------------------------------
    def state_dict(self):
        return OrderedDict((n, t.detach().clone()) for n, t in self.params.items())

    def load_state_dict(self, sd):
        for n in self.params:
            self.params[n] = sd[n].detach().to(self.dtype).clone()

    def to(self, dtype):
        self.dtype = dtype
        for n in self.params:
            self.params[n] = self.params[n].detach().to(dtype)
        if self.acc is not None:
            self.acc = {n: t.to(dtype) for n, t in self.acc.items()}
        return self


def step_decay(initial_lr, decay_factor, step_size, epoch):
    """schedule.py:17-19"""
    return initial_lr * (decay_factor ** np.floor(epoch / step_size))
