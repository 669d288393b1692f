"""Functional layer blocks with the keyword names of the reference (mvae/layer_blocks.py), executed eagerly on B200.

The reference functions build Keras graph nodes; called on arrays (as its tests do, tests/test_layer_blocks.py) they
create freshly initialised weights and run.  These mirrors do the same: numpy in -> numpy out, CUDA tensor in ->
CUDA tensor out, every arithmetic step a libmvae_b200.so kernel.  Covered: gaussian_kernel (:980-1002),
gaussian_filter_block (:1008-1050), laplacian_transform_split/merge (:23-185, trainable=False), squeeze_excite_block
(:418-462), mobilenetV3_block (:556-648), basic_block (:893-974).  The reference's other blocks are never reached from
mvae/multiscale_vae.py and are out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import engine as _E
from ._lib import check

DEFAULT_DROPOUT_RATIO = 0.0
DEFAULT_CHANNEL_INDEX = 3
DEFAULT_KERNEL_REGULARIZER = "l1"
DEFAULT_KERNEL_INITIALIZER = "glorot_normal"
DEFAULT_GAUSSIAN_XY_MAX = (1, 1)
DEFAULT_GAUSSIAN_KERNEL_SIZE = (3, 3)

gaussian_kernel = _E.gaussian_kernel


def _to_dev(x):
    was_numpy = not torch.is_tensor(x)
    t = torch.as_tensor(np.asarray(x, dtype=np.float32)) if was_numpy else x
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    _lib.require_b200(dev.index or 0)
    return t.to(dev, torch.float32).contiguous(), was_numpy, dev


def _ret(y, was_numpy):
    return y.cpu().numpy() if was_numpy else y


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def gaussian_filter_block(input_layer, kernel_size=DEFAULT_GAUSSIAN_KERNEL_SIZE, strides=(1, 1), dilation_rate=(1, 1),
                          padding="same", xy_max=DEFAULT_GAUSSIAN_XY_MAX):
    """Frozen depthwise Gaussian, SAME zero padding (layer_blocks.py:1008-1050)."""
    if tuple(strides) != (1, 1) or tuple(dilation_rate) != (1, 1) or padding != "same":
        raise NotImplementedError("only strides (1,1), dilation (1,1), padding 'same' are on the hot path")
    x, was_numpy, dev = _to_dev(input_layer)
    B, H, W, Cc = x.shape
    k = gaussian_kernel(kernel_size, xy_max).astype(np.float32)
    taps = (C.c_float * k.size)(*[float(v) for v in k.ravel()])
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        check(_lib.load().mvae_gaussian_filter(x.data_ptr(), y.data_ptr(), B, H, W, Cc, taps, k.shape[0], k.shape[1],
                                               _stream(dev)), "gaussian_filter")
    return _ret(y, was_numpy)


class _PyramidModel:
    def __init__(self, fn, name):
        self._fn, self.name = fn, name

    def __call__(self, x):
        return self._fn(x)

    predict = __call__


def laplacian_transform_split(input_dims, levels: int, name: str = None, min_value: float = 0.0,
                              max_value: float = 255.0, gaussian_xy_max: tuple = DEFAULT_GAUSSIAN_XY_MAX,
                              gaussian_kernel_size: tuple = DEFAULT_GAUSSIAN_KERNEL_SIZE, diff_mode="laplacian"):
    """Normalise and split into a Laplacian pyramid (layer_blocks.py:23-101).  Returns a callable model."""
    H, W, Cc = input_dims
    k = gaussian_kernel(gaussian_kernel_size, gaussian_xy_max).astype(np.float32)
    taps = (C.c_float * k.size)(*[float(v) for v in k.ravel()])
    mode = {"laplacian": _lib.DIFF_LAPLACIAN, "no_upsample": _lib.DIFF_NO_UPSAMPLE}[diff_mode]

    def run(x):
        x, was_numpy, dev = _to_dev(x)
        B = x.shape[0]
        if tuple(x.shape[1:]) != (H, W, Cc):
            raise ValueError(f"expected input {(H, W, Cc)}, got {tuple(x.shape[1:])}")
        lib = _lib.load()
        bands = [torch.empty((B, H >> i, W >> i, Cc), dtype=torch.float32, device=dev) for i in range(levels)]
        ptrs = (C.c_void_p * levels)(*[b.data_ptr() for b in bands])
        ws = torch.empty(lib.mvae_pyramid_split_workspace_bytes(B, H, W, Cc, levels) // 4 + 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mvae_pyramid_split(x.data_ptr(), ptrs, ws.data_ptr(), B, H, W, Cc, levels, float(min_value),
                                         float(max_value), taps, k.shape[0], k.shape[1], mode, _stream(dev)), "pyramid_split")
        return [_ret(b, was_numpy) for b in bands]

    return _PyramidModel(run, name)


def laplacian_transform_merge(input_dims, levels: int, name: str = None, min_value: float = 0.0, max_value: float = 255.0,
                              trainable: bool = False, filters: int = 32, activation: str = "relu",
                              kernel_regularizer: str = DEFAULT_KERNEL_REGULARIZER,
                              kernel_initializer: str = DEFAULT_KERNEL_INITIALIZER):
    """Merge Laplacian pyramid stages and denormalise (layer_blocks.py:107-185, trainable=False branch)."""
    if trainable:
        raise NotImplementedError("the trainable merge (layer_blocks.py:146-169) is not used by MultiscaleVAE")
    H, W, Cc = input_dims[0]

    def run(xs):
        conv = [_to_dev(x) for x in xs]
        ts, was_numpy, dev = [c[0] for c in conv], conv[0][1], conv[0][2]
        B = ts[0].shape[0]
        lib = _lib.load()
        ptrs = (C.c_void_p * levels)(*[t.data_ptr() for t in ts])
        r0 = torch.empty((B, H, W, Cc), dtype=torch.float32, device=dev)
        out = torch.empty_like(r0)
        ws = torch.empty(lib.mvae_pyramid_merge_workspace_bytes(B, H, W, Cc, levels) // 4 + 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mvae_pyramid_merge_fwd(ptrs, r0.data_ptr(), ws.data_ptr(), B, H, W, Cc, levels, _stream(dev)), "merge")
            check(lib.mvae_denormalize_clip(r0.data_ptr(), out.data_ptr(), r0.numel(), float(min_value), float(max_value),
                                            _stream(dev)), "denormalize")
        return _ret(out, was_numpy)

    return _PyramidModel(run, name)


# ------------------------------------------------------------------------------------------------------------------
# Blocks with weights: a throw-away parameter store + the same op classes the model engine uses
# ------------------------------------------------------------------------------------------------------------------
class _Mini:
    """Just enough of engine.Engine for the op classes: allocation, parameter pointers, stream."""

    def __init__(self, dev, ps, B):
        self.lib, self.device, self.ps, self.B = _lib.load(), dev, ps, B
        self.training, self.precision = False, _lib.PREC_FP32
        self.s = _stream(dev)
        self._z = []

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def zeros(self, n, bwd=False):
        t = torch.zeros(max(n, 1), dtype=torch.float32, device=self.device)
        z = _E._Z(0, n)
        z.ptr = t.data_ptr()
        self._z.append(t)
        return z

    def new_T(self, shape, act):
        return _E.T(self.empty(shape), None, act)

    def side(self, fn):
        fn()

    def wgrad(self, desc, x, gate, dy, dw, db):      # inline (the engine defers and batches them in graph mode)
        import ctypes
        from multiscale_variational_autoencoder_b200._lib import check
        check(self.lib.mvae_conv2d_wgrad(ctypes.byref(desc), x, gate, dy, dw, db, self.s), "conv2d_wgrad")


_seed = [0]


def _fresh_store(dev):
    _seed[0] += 1
    return _E.ParamStore(dev, seed=1000 + _seed[0])


def mobilenetV3_block(input_layer, filters: int = 32, squeeze_units: int = -1, activation: str = "relu",
                      dropout_ratio: float = None, use_batchnorm: bool = False, prefix: str = "mobilenetV3_",
                      regularizer: str = DEFAULT_KERNEL_REGULARIZER, initializer: str = DEFAULT_KERNEL_INITIALIZER,
                      channels_index: int = DEFAULT_CHANNEL_INDEX, training: bool = False):
    """layer_blocks.py:556-648 with fresh glorot_normal weights."""
    if input_layer is None:
        raise ValueError("input_layer cannot be empty")
    if filters <= 0:
        raise ValueError("Filters should be > 0")
    if dropout_ratio is not None and (dropout_ratio > 1.0 or dropout_ratio < 0.0):
        raise ValueError("Dropout ration must be [0, 1]")
    if activation != "relu" or (squeeze_units not in (-1, None, filters)):
        raise NotImplementedError("hot-path configuration only: relu, squeeze_units == filters")
    x, was_numpy, dev = _to_dev(input_layer)
    with torch.cuda.device(dev):
        ps = _fresh_store(dev)
        _E.declare_mbv3(ps, prefix, x.shape[3], filters)
        ps.finalize()
        m = _Mini(dev, ps, x.shape[0])
        m.training = training
        op = _E.MobileNetV3(m, _E.T(x), prefix, filters)
        op.fwd()
    return _ret(op.y.data, was_numpy)


def squeeze_excite_block(input_layer, squeeze_units: int = -1, use_batchnorm: bool = False, prefix="squeeze_excite_",
                         initializer=DEFAULT_KERNEL_INITIALIZER, regularizer=DEFAULT_KERNEL_REGULARIZER,
                         channels_index: int = DEFAULT_CHANNEL_INDEX, training: bool = False):
    """layer_blocks.py:418-462: GAP -> Dense relu -> [BatchNorm] -> Dense hard_sigmoid -> Multiply."""
    if input_layer is None:
        raise ValueError("input_layer cannot be empty")
    if np.ndim(input_layer) != 4:
        raise ValueError("works only on 4d tensors")
    x, was_numpy, dev = _to_dev(input_layer)
    B, H, W, Cc = x.shape
    if squeeze_units not in (-1, None, Cc) and squeeze_units > 0:
        raise NotImplementedError("hot-path configuration only: squeeze_units == channels")
    with torch.cuda.device(dev):
        ps = _fresh_store(dev)
        _E.declare_dense(ps, prefix + "dense0", Cc, Cc, _lib.REG_L1)
        _E.declare_bn(ps, prefix + "batchnorm0", Cc)
        _E.declare_dense(ps, prefix + "dense1", Cc, Cc, _lib.REG_L1)
        ps.finalize()
        lib, s = _lib.load(), _stream(dev)
        gap = torch.zeros((B, Cc), dtype=torch.float32, device=dev)
        gate = torch.empty((B, Cc), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.mvae_se_gate_ws_floats(B, Cc), dtype=torch.float32, device=dev)
        y = torch.empty_like(x)
        P = lambda n: ps.ptr(prefix + n)
        check(lib.mvae_se_dgate_reduce(x.data_ptr(), 0, gap.data_ptr(), B, H * W, Cc, s), "gap")
        # without batchnorm the BN parameters stay at identity (gamma 1, beta 0, moving stats 0/1, eps 0)
        check(lib.mvae_se_gate_fwd(gap.data_ptr(), P("dense0/kernel"), P("dense0/bias"), P("batchnorm0/gamma"),
                                   P("batchnorm0/beta"), P("dense1/kernel"), P("dense1/bias"), P("batchnorm0/moving_mean"),
                                   P("batchnorm0/moving_variance"), gate.data_ptr(), ws.data_ptr(), B, Cc, H * W,
                                   _E.SE_BN_EPS if use_batchnorm else 0.0, _E.SE_BN_MOM,
                                   1 if (training and use_batchnorm) else 0, s), "se_gate")
        check(lib.mvae_channel_scale(x.data_ptr(), gate.data_ptr(), y.data_ptr(), B, H * W, Cc, s), "channel_scale")
    return _ret(y, was_numpy)


def basic_block(input_layer, block_type="encoder", filters=[64], kernel_size=[(3, 3)], strides=[(1, 1)],
                initializer: str = DEFAULT_KERNEL_INITIALIZER, regularizer: str = DEFAULT_KERNEL_REGULARIZER,
                use_batchnorm: bool = False, use_dropout: bool = False, prefix: str = "block_", training: bool = False):
    """layer_blocks.py:893-974: per entry an optional strided Conv2D / Conv2DTranspose, then a mobilenetV3 block."""
    if len(filters) != len(kernel_size) or len(filters) != len(strides) or len(filters) <= 0:
        raise ValueError("len(filters) [{0}] should be equal to len(kernel_size) [{1}] and len(strides) [{2}]".format(
            len(filters), len(kernel_size), len(strides)))
    if block_type != "encoder" and block_type != "decoder":
        raise ValueError("block_type should be encoder or decoder")
    if use_batchnorm or use_dropout:
        raise NotImplementedError("MultiscaleVAE calls basic_block with use_batchnorm=False, use_dropout=False")
    x, was_numpy, dev = _to_dev(input_layer)
    with torch.cuda.device(dev):
        ps = _fresh_store(dev)
        prev = x.shape[3]
        plan = []
        for i in range(len(filters)):
            f, k, s = int(filters[i]), tuple(kernel_size[i]), tuple(strides[i])
            pi = f"{prefix}_{i}_"
            if s != (1, 1) or f != prev:
                name = pi + ("conv" if block_type == "encoder" else "conv_transpose")
                _E.declare_conv(ps, name, k[0], k[1], prev, f, _lib.REG_L1, transpose=(block_type == "decoder"))
                plan.append((block_type, name, k, s, f))
            _E.declare_mbv3(ps, pi + "mobilenetV3_", f, f)
            plan.append(("mbv3", pi + "mobilenetV3_", None, None, f))
            prev = f
        ps.finalize()
        m = _Mini(dev, ps, x.shape[0])
        m.training = training
        t = _E.T(x)
        for kind, name, k, s, f in plan:
            if kind == "encoder":
                op = _E.Conv2D(m, t, name + "/kernel", name + "/bias", k[0], k[1], s, f)
            elif kind == "decoder":
                op = _E.Conv2DTranspose(m, t, name + "/kernel", name + "/bias", k[0], k[1], s, f)
            else:
                op = _E.MobileNetV3(m, t, name, f)
            op.fwd()
            t = op.y
    return _ret(t.data, was_numpy)
