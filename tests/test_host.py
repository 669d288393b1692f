"""Host-side logic that needs no GPU: plan/parameter layout, argument validation, schedule, failure without CUDA."""
import numpy as np
import pytest
import torch

from multiscale_variational_autoencoder_b200 import _lib, engine, schedule
from oracle import mvae_oracle as O

CFG1 = dict(input_dims=(32, 32, 3), z_dims=[128, 64, 32],
            encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (2, 2), (1, 1)]})


def _store(cfg, coord=None):
    enc = cfg["encoder"]
    dec = {k: enc[k][::-1] for k in ("filters", "strides", "kernel_size")}
    spec = engine.Spec(cfg["input_dims"], cfg["z_dims"], enc, dec, 0.0, 255.0, 0.5, coord, 1.0, "no_upsample")
    ps = engine.ParamStore(torch.device("cpu"), seed=7)
    spec.declare_params(ps)
    ps.finalize()
    return spec, ps


def test_product_gaussian_kernel_equals_reference_output_bitwise():
    """engine.gaussian_kernel (== layer_blocks.gaussian_kernel of the product) against the reference function's output
    (layer_blocks.py:980-1002 executed by tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gaussian_kernel.npz"))
    for key in g.files:
        size, nsig = key.split("_")
        size = tuple(int(v) for v in size[1:].split("x"))
        nsig = tuple(int(v) for v in nsig[1:].split("x"))
        assert np.array_equal(engine.gaussian_kernel(size, nsig), g[key]), key


def test_param_layout_matches_oracle_names_and_counts():
    spec, ps = _store(CFG1)
    sd = ps.state_dict()
    m = O.OracleMVAE(CFG1["input_dims"], CFG1["z_dims"], CFG1["encoder"])
    assert sorted(sd.keys()) == sorted(m.params.keys())      # same Keras variable names
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(m.params[k].shape), k
    assert sum(v.numel() for v in sd.values()) == 1_098_601
    trainable = sum(s[1] for s in ps.segs)
    assert trainable == 1_097_257
    regs = {s[5]: s[4] for s in ps.segs}
    for k in m.params:
        if m.trainable[k]:
            assert regs[k] == m.reg[k], k


def test_fused_mu_logvar_segments_are_column_halves():
    spec, ps = _store(CFG1)
    seg = {s[5]: s for s in ps.segs}
    off, count, width, ld = seg["encoder_0_mu/kernel"][:4]
    off2 = seg["encoder_0_log_var/kernel"][0]
    assert (count, width, ld) == (2048 * 128, 128, 256) and off2 == off + 128
    mu = ps.get("encoder_0_mu/kernel")
    lv = ps.get("encoder_0_log_var/kernel")
    fused = ps.view("encoder_0_mu_log_var/kernel")
    assert torch.equal(fused[:, :128], mu) and torch.equal(fused[:, 128:], lv)
    assert ps.get("encoder_0_mu/bias").shape == (128,)


def test_state_dict_roundtrip():
    spec, ps = _store(CFG1)
    sd = ps.state_dict()
    sd2 = {k: torch.randn_like(v) for k, v in sd.items()}
    ps.load_state_dict(sd2)
    for k, v in ps.state_dict().items():
        assert torch.equal(v, sd2[k]), k


def test_glorot_normal_statistics():
    spec, ps = _store(CFG1)
    w = ps.get("decoder_0_dense/kernel")
    std = np.sqrt(2.0 / (128 + 2048))
    assert abs(float(w.std()) - std) / std < 0.02
    assert float(w.abs().max()) <= 2 * std / 0.87962566 + 1e-6
    assert float(ps.get("decoder_0_batchnorm/gamma").min()) == 1.0
    assert float(ps.get("decoder_0_batchnorm/moving_variance").min()) == 1.0
    assert float(ps.get("encoder_0_conv_base/bias").abs().max()) == 0.0


def test_validation_errors():
    enc = CFG1["encoder"]
    dec = {k: enc[k][::-1] for k in ("filters", "strides", "kernel_size")}
    with pytest.raises(ValueError):          # 5 levels with total stride 4: SURVEY App. C-7
        s = engine.Spec((32, 32, 3), [8] * 5, enc, dec, 0, 255, .5, None, 1.0, "no_upsample")
        s.declare_params(engine.ParamStore(torch.device("cpu")))
    with pytest.raises(ValueError):          # odd scale
        engine.Spec((36, 36, 3), [8] * 4, enc, dec, 0, 255, .5, None, 1.0, "no_upsample")
    with pytest.raises(ValueError):          # one level
        engine.Spec((32, 32, 3), [8], enc, dec, 0, 255, .5, None, 1.0, "no_upsample")
    with pytest.raises(ValueError):          # list-length mismatch (layer_blocks.py:918-924)
        engine.Spec.entries({"filters": [32, 32], "kernel_size": [(3, 3)], "strides": [(1, 1)]})


def test_coord_conv_widens_conv_base():
    spec, ps = _store(CFG1, coord="xyr")
    assert ps.entries["encoder_0_conv_base/kernel"]["shape"] == (3, 3, 6, 32)


def test_schedule_matches_reference_formula():
    f = schedule.step_decay_schedule(0.01, 0.75, 20)
    assert f(0) == 0.01 and abs(f(20) - 0.0075) < 1e-12 and abs(f(45) - 0.01 * 0.75 ** 2) < 1e-12
    assert f(45) == O.step_decay(0.01, 0.75, 20, 45)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the failure mode without a GPU")
def test_product_fails_loudly_without_gpu():
    from multiscale_variational_autoencoder_b200 import MultiscaleVAE
    with pytest.raises(_lib.MvaeError):
        MultiscaleVAE(**CFG1)


def test_bench_accounting_of_algorithmic_bytes_and_flops():
    """bench.py::account() -- the algorithmic byte / FLOP counts behind `roofline.achieved`."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from multiscale_variational_autoencoder_b200 import _lib
    from multiscale_variational_autoencoder_b200.engine import _descs
    d = _lib.ConvDesc(256, 16, 16, 32, 1, 1, 1, 1, 32, 0, 1)
    key, by, fl = bench.account("mvae_conv2d_wgrad", (C.byref(d), 1, 0, 1, 1, 1, 0))
    n = 256 * 16 * 16 * 32
    assert key == "B256 16x16x32->32 k1 s1"
    assert by == 4.0 * (2 * n + 32 * 32) and fl == 2.0 * 256 * 16 * 16 * 32 * 32
    # forward with a residual reads one more output-sized tensor
    _, by_res, _ = bench.account("mvae_conv2d_fwd", (C.byref(d), 1, 1, 1, 0, 1, 0, 1, 0))
    assert by_res == by + 4.0 * n
    # strided 3x3: output is a quarter of the input pixels
    d2 = _lib.ConvDesc(256, 32, 32, 32, 3, 3, 2, 2, 32, 0, 1)
    _, by2, fl2 = bench.account("mvae_conv2d_wgrad", (C.byref(d2), 1, 0, 1, 1, 1, 0))
    assert by2 == 4.0 * (256 * 32 * 32 * 32 + 256 * 16 * 16 * 32 + 9 * 32 * 32)
    assert fl2 == 2.0 * 256 * 16 * 16 * 9 * 32 * 32
    # a batched launch is the sum of its members
    arr = _descs([d, d, d2])
    key_b, by_b, fl_b = bench.account("mvae_conv2d_wgrad_batched", (3, arr, None, None, None, None, None, 0))
    assert by_b == 2 * by + by2 and fl_b == 2 * fl + fl2 and "2@16x16" in key_b and "1@32x32" in key_b
    # Dense: x, W and y once
    key_d, by_d, fl_d = bench.account("mvae_dense_fwd", (256, 8192, 256, 1, 1, 1, 0, 1, 1, 0, 1, 0))
    assert by_d == 4.0 * (256 * 8192 + 8192 * 256 + 256 * 256) and fl_d == 2.0 * 256 * 8192 * 256
    assert key_d == "M256 K8192 N256"


def test_visualisation_callback_writes_the_three_collages(tmp_path):
    """callbacks.SaveIntermediateResultsCallback (mvae/callbacks.py:16-138) on a stand-in VAE: file names, cadence,
    interpolation rows, collage layout, nearest-neighbour resize."""
    from PIL import Image
    from multiscale_variational_autoencoder_b200 import callbacks

    class FakeVAE:
        def __init__(self):
            self.model_encode = self.model_decode = self
            self.calls = []

        def predict(self, v):
            v = np.asarray(v)
            self.calls.append(v.shape)
            if v.ndim == 4:                                   # encode: mean colour per image, tiled to 6 latents
                return np.tile(v.mean(axis=(1, 2)), (1, 2)).astype(np.float32)
            return np.broadcast_to(v[:, None, None, :3], (v.shape[0], 4, 4, 3)).astype(np.float32)

        @staticmethod
        def normalize(v):
            return v / 255.0

    imgs = np.stack([np.full((4, 4, 3), 10.0 * i, dtype=np.float32) for i in range(16)])
    vae = FakeVAE()
    cb = callbacks.SaveIntermediateResultsCallback(str(tmp_path), 5, 0, imgs, vae, resize_shape=(32, 32))
    cb.on_epoch_begin(0)
    assert cb.on_batch_end(3) == [] and vae.calls == []                   # only every 5th batch
    files = cb.on_batch_end(10)
    assert [f.split("/")[-1] for f in files] == ["img_001_10.png", "samples_001_10.png", "interpolations_001_10.png"]
    im = np.asarray(Image.open(files[0]))
    assert im.shape == (32, 32, 3)
    # 16 images of 4x4 -> 4x4 grid -> 16x16 -> nearest x2: cell (r, q) is image 4r+q, value 10*(4r+q)
    assert im[0, 0, 0] == 0 and im[0, 8, 0] == 10 and im[8, 0, 0] == 40 and im[31, 31, 0] == 150
    enc = np.arange(16 * 6, dtype=np.float32).reshape(16, 6)
    it = cb.interpolations(enc)
    assert np.allclose(it[0], enc[0]) and np.allclose(it[3], enc[1]) and np.allclose(it[5], enc[1] * (2 / 3) + enc[2] / 3)
    g = callbacks.collage(np.ones((5, 2, 3, 1)))
    assert g.shape == (4, 9, 1) and g[:2].sum() == 18 and g[2:, 6:].sum() == 0
    assert np.array_equal(callbacks.resize_nearest(np.arange(4).reshape(2, 2), (4, 4))[0], [0, 0, 1, 1])
