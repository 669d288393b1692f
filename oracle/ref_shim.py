"""Execute the PURE-NUMPY parts of the reference in this container.  TEST INFRASTRUCTURE ONLY.

TensorFlow/Keras cannot be installed here, so the reference package does not import.  Two pieces of
it are nevertheless plain Python once `keras` resolves:

  * `mvae/layer_blocks.py:980-1002`  gaussian_kernel  (numpy only)
  * `mvae/coord.py:88-133`           _CoordinateChannel.call, rank 2 (a dozen Keras-backend calls,
                                     each with an unambiguous numpy meaning)
  * `mvae/multiscale_vae.py:437-504` compile(): its loss closures (vae_r_loss, vae_r_experimental_loss, vae_kl_loss,
                                     vae_loss) are K.abs / K.mean / K.square / K.exp / K.sum and slicing only; calling
                                     the unbound method on a stub whose `_model_trainable.compile` records its
                                     arguments hands them out (`load_reference_losses`)
  * `mvae/schedule.py:7-21`          step_decay_schedule (numpy only; the LearningRateScheduler stub keeps the function)
  * `mvae/multiscale_vae.py:508-557` train(): the arguments it passes to `fit` and the callbacks it builds

This module installs a minimal stand-in `keras` (numpy-backed `keras.backend`) into `sys.modules`,
loads those two reference files *from where they lie* under /root/reference (nothing is copied) and
exposes the reference callables.  `tests/golden/make_golden.py` uses it to write golden vectors; the
GPU box never needs /root/reference because the vectors are committed.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MVAE_REFERENCE_ROOT", "/root/reference")


def _numpy_backend():
    K = types.ModuleType("keras.backend")
    K.floatx = lambda: "float32"
    K.image_data_format = lambda: "channels_last"
    K.shape = lambda x: np.asarray(np.shape(x))
    K.int_shape = lambda x: tuple(np.shape(x))
    K.stack = lambda xs, axis=0: np.stack([np.asarray(v) for v in xs], axis=axis)
    K.ones = lambda shape, dtype="float32": np.ones(tuple(int(s) for s in np.asarray(shape).ravel()), dtype=dtype)
    K.arange = lambda start, stop=None, step=1, dtype="int32": np.arange(start, stop, step, dtype=dtype)
    K.expand_dims = lambda x, axis=-1: np.expand_dims(x, axis)
    K.tile = lambda x, n: np.tile(x, tuple(int(s) for s in np.asarray(n).ravel()))
    K.permute_dimensions = lambda x, pattern: np.transpose(x, pattern)
    K.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    K.concatenate = lambda xs, axis=-1: np.concatenate(xs, axis=axis)
    K.sqrt = np.sqrt
    K.square = np.square
    K.abs = np.abs
    K.exp = np.exp
    K.mean = lambda x, axis=None, keepdims=False: np.mean(x, axis=tuple(axis) if isinstance(axis, list) else axis,
                                                          keepdims=keepdims)
    K.sum = lambda x, axis=None, keepdims=False: np.sum(x, axis=tuple(axis) if isinstance(axis, list) else axis,
                                                        keepdims=keepdims)

    def batch_dot(x, y, axes):
        # 3-D case used by coord.py: contract x's axis axes[0] with y's axis axes[1], batch on axis 0
        assert x.ndim == 3 and y.ndim == 3
        xs = "bik" if axes[0] == 2 else "bki"
        ys = "bkj" if axes[1] == 1 else "bjk"
        return np.einsum(f"{xs},{ys}->bij", x, y)

    K.batch_dot = batch_dot
    return K


def install_keras_shim():
    if "keras" in sys.modules and not getattr(sys.modules["keras"], "_mvae_shim", False):
        raise RuntimeError("a real keras is importable; the shim is not needed")
    keras = types.ModuleType("keras")
    keras._mvae_shim = True
    K = _numpy_backend()
    layers = types.ModuleType("keras.layers")

    class Layer:
        def __init__(self, **kwargs):
            self.built = False

        def __call__(self, inputs, **kwargs):
            if not self.built:
                self.build(np.shape(inputs))
            return self.call(inputs, **kwargs)

        def get_config(self):
            return {}

    class InputSpec:
        def __init__(self, **kwargs):
            self.__dict__.update(kwargs)

    layers.Layer, layers.InputSpec = Layer, InputSpec
    utils = types.ModuleType("keras.utils")
    _custom = {}
    utils.get_custom_objects = lambda: _custom
    keras.backend, keras.layers, keras.utils = K, layers, utils

    # recording stand-ins for what compile() / train() / schedule.py construct
    class _Record:
        def __init__(self, *args, **kwargs):
            self.args, self.kwargs = args, kwargs

    optimizers = types.ModuleType("keras.optimizers")
    optimizers.Adagrad = type("Adagrad", (_Record,), {})
    cb = types.ModuleType("keras.callbacks")
    cb.LearningRateScheduler = type("LearningRateScheduler", (_Record,), {"schedule": property(lambda self: self.args[0])})
    cb.ModelCheckpoint = type("ModelCheckpoint", (_Record,), {})
    cb.Callback = type("Callback", (), {})
    keras.optimizers, keras.callbacks = optimizers, cb
    sys.modules.update({"keras": keras, "keras.backend": K, "keras.layers": layers, "keras.utils": utils,
                        "keras.optimizers": optimizers, "keras.callbacks": cb})
    return keras


def _load(relpath, modname, package=None):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (layer_blocks_module, coord_module) of the reference, running on the shim."""
    install_keras_shim()
    pkg = types.ModuleType("mvae_ref")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "mvae")]
    sys.modules["mvae_ref"] = pkg
    _load("mvae/custom_logger.py", "mvae_ref.custom_logger", "mvae_ref")
    lb = _load("mvae/layer_blocks.py", "mvae_ref.layer_blocks", "mvae_ref")
    coord = _load("mvae/coord.py", "mvae_ref.coord", "mvae_ref")
    return lb, coord


class _Recorder:
    """Stands in for the keras.Model objects compile() / train() call into: keeps the arguments."""

    def __init__(self):
        self.compiled, self.fitted = None, None

    def compile(self, **kwargs):
        self.compiled = kwargs

    def fit(self, *args, **kwargs):
        self.fitted = (args, kwargs)


def load_reference_model_module():
    """mvae/multiscale_vae.py and mvae/schedule.py of the reference on the shim (callbacks.py needs matplotlib / skimage /
    a missing `collage` helper, SURVEY App. C, and is replaced by a recording stub)."""
    install_keras_shim()
    load_reference()
    cbs = types.ModuleType("mvae_ref.callbacks")

    class SaveIntermediateResultsCallback:
        def __init__(self, *args):
            self.args = args

    cbs.SaveIntermediateResultsCallback = SaveIntermediateResultsCallback
    sys.modules["mvae_ref.callbacks"] = cbs
    sys.modules["mvae_ref"].callbacks = cbs
    sched = _load("mvae/schedule.py", "mvae_ref.schedule", "mvae_ref")
    sys.modules["mvae_ref"].schedule = sched
    sys.modules["mvae_ref"].layer_blocks = sys.modules["mvae_ref.layer_blocks"]
    mv = _load("mvae/multiscale_vae.py", "mvae_ref.multiscale_vae", "mvae_ref")
    return mv, sched


def load_reference_losses(input_dims, mu, log_var, r_loss_factor, kl_loss_factor, learning_rate=0.01, clip_norm=1.0):
    """Runs the reference's MultiscaleVAE.compile (multiscale_vae.py:437-504) on a stub instance and returns what it
    passed to keras: dict(loss=vae_loss, metrics=[vae_r_loss, vae_kl_loss], optimizer=<recorded Adagrad>).  The
    closures are the reference's own code objects, evaluated with numpy arrays."""
    mv, _ = load_reference_model_module()
    stub = types.SimpleNamespace(_inputs_dims=tuple(input_dims), _mu=mu, _log_var=log_var, _model_trainable=_Recorder(),
                                 learning_rate=None)
    mv.MultiscaleVAE.compile(stub, learning_rate, r_loss_factor, kl_loss_factor, clip_norm)
    out = dict(stub._model_trainable.compiled)
    # vae_r_experimental_loss is only reachable as a closure cell of vae_loss
    cells = dict(zip(out["loss"].__code__.co_freevars, (c.cell_contents for c in out["loss"].__closure__)))
    out["vae_r_experimental_loss"] = cells["vae_r_experimental_loss"]
    return out


def run_reference_train(x_train, batch_size, epochs, run_folder, **kw):
    """Runs the reference's MultiscaleVAE.train (multiscale_vae.py:508-557) on a stub and returns the recorded
    `fit(*args, **kwargs)` call."""
    mv, _ = load_reference_model_module()
    stub = types.SimpleNamespace(_learning_rate=kw.pop("learning_rate", 0.01), _model_trainable=_Recorder())
    mv.MultiscaleVAE.train(stub, x_train, batch_size, epochs, run_folder, **kw)
    return stub._model_trainable.fitted
