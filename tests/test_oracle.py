"""CPU tests of the oracle: golden vectors produced by executing the reference (tests/golden/make_golden.py), the
reference's own known-answer tests (tests/test_layer_blocks.py in the reference), and self-consistency."""
import os

import numpy as np
import pytest
import torch

from oracle import mvae_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_gaussian_kernel_matches_reference_function():
    g = np.load(os.path.join(GOLD, "gaussian_kernel.npz"))
    for key in g.files:
        size, nsig = key.split("_")
        size = tuple(int(v) for v in size[1:].split("x"))
        nsig = tuple(int(v) for v in nsig[1:].split("x"))
        assert np.array_equal(O.gaussian_kernel(size, nsig), g[key]), key
    # SURVEY App. A.1 constants
    k = O.gaussian_kernel((3, 3), (2, 2)).astype(np.float32)
    assert np.allclose([k[1, 1], k[0, 1], k[0, 0]], [0.61934704, 0.08381951, 0.011343736], rtol=1e-6)


def test_coord_channels_match_reference_layer():
    c = np.load(os.path.join(GOLD, "coord_channels.npz"))
    for n in "abc":
        y = O.coordinate_channels_2d(torch.from_numpy(c[n + "_x"]), bool(c[n + "_r"])).numpy()
        assert y.shape == c[n + "_y"].shape
        assert np.abs(y - c[n + "_y"]).max() <= 1e-6


# ---- the reference's four value-asserting tests (tests/test_layer_blocks.py:9-39,160-190) ---------------------
def test_gaussian_filter_block_all_zeros():
    x = torch.zeros(3, 256, 256, 3)
    y = O.gaussian_filter(x)
    assert tuple(y.shape) == (3, 256, 256, 3) and bool((y == 0).all())


def test_gaussian_filter_block_one_channel_all_ones():
    y = O.gaussian_filter(torch.ones(3, 16, 16, 1))
    assert tuple(y.shape) == (3, 16, 16, 1) and bool((y[:, 1:15, 1:15, :] == 1.0).all())


def test_gaussian_filter_block_three_channels_all_ones():
    y = O.gaussian_filter(torch.ones(3, 9, 9, 7))
    assert tuple(y.shape) == (3, 9, 9, 7) and bool((y[:, 1:8, 1:8, :] == 1.0).all())


def test_laplacian_transform_split_merge():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(18, 32, 32, 3, generator=g) * 255
    bands = O.pyramid_split(x, 3, 0.0, 255.0, nsig=(1, 1), mode="laplacian")
    assert [tuple(b.shape) for b in bands] == [(18, 32, 32, 3), (18, 16, 16, 3), (18, 8, 8, 3)]
    m = O.pyramid_merge(bands, 0.0, 255.0)
    assert tuple(m.shape) == (18, 32, 32, 3)
    assert bool(((m - x).abs()[:, 1:31, 1:31, :] <= 1e-3).all())


# ---- TF conventions restated by hand -------------------------------------------------------------
def test_same_padding_stride2_is_asymmetric():
    # k=3, s=2, even size: pad (0 before, 1 after)
    assert O.same_pads(32, 3, 2) == (0, 1, 16)
    assert O.same_pads(32, 5, 2) == (1, 2, 16)
    assert O.same_pads(32, 3, 1) == (1, 1, 32)
    x = torch.zeros(1, 4, 4, 1)
    x[0, 0, 0, 0] = 1.0
    w = torch.zeros(3, 3, 1, 1)
    w[0, 0, 0, 0] = 1.0          # tap (ky=0,kx=0) reads x[2*oy + 0 - 0]
    y = O.conv2d_same(x, w, None, (2, 2))
    assert y[0, 0, 0, 0] == 1.0 and y.sum() == 1.0


def test_conv_transpose_is_adjoint_of_same_conv():
    g = torch.Generator().manual_seed(1)
    for k, s, n in [(3, 2, 4), (3, 1, 5), (5, 2, 3)]:
        w = torch.randn(k, k, 2, 3, generator=g, dtype=torch.float64)      # forward conv 2 -> 3 channels
        x = torch.randn(1, n * s, n * s, 2, generator=g, dtype=torch.float64)
        y = torch.randn(1, n, n, 3, generator=g, dtype=torch.float64)
        lhs = (O.conv2d_same(x, w, None, (s, s)) * y).sum()
        rhs = (x * O.conv2d_transpose_same(y, w, None, (s, s))).sum()       # transpose kernel layout (kh,kw,Cout,Cin)
        assert abs(float(lhs - rhs)) < 1e-9 * max(1.0, abs(float(lhs)))


def test_bilinear_up2_half_pixel():
    x = torch.tensor([0.0, 4.0]).view(1, 1, 2, 1)
    y = O.bilinear_up2(x.expand(1, 2, 2, 1).contiguous())[0, 0, :, 0]
    assert torch.allclose(y, torch.tensor([0.0, 1.0, 3.0, 4.0]))


def test_hard_sigmoid_and_loss_crop():
    assert torch.allclose(O.hard_sigmoid(torch.tensor([-3.0, 0.0, 1.0, 3.0])), torch.tensor([0.0, 0.5, 0.7, 1.0]))
    m = O.OracleMVAE((32, 32, 3), [4, 4], encoder={"filters": [8], "kernel_size": [(3, 3)], "strides": [(1, 1)]})
    y = torch.zeros(1, 32, 32, 3)
    yh = torch.zeros(1, 32, 32, 3)
    yh[:, 8:24, 8:24, :] = 1.0          # exactly the centre crop [8:24]
    r = m.r_loss(y, yh)
    assert abs(float(r) - (0.25 + 0.5 * (0.25 + 1.0))) < 1e-6


def test_param_counts_config1():
    m = O.OracleMVAE((32, 32, 3), [128, 64, 32],
                     encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (2, 2), (1, 1)]})
    assert sum(t.numel() for n, t in m.params.items() if m.trainable[n]) == 1_097_257
    assert sum(t.numel() for t in m.params.values()) == 1_098_601


def test_stride_constraint_raises():
    with pytest.raises(ValueError):
        O.OracleMVAE((32, 32, 3), [8] * 5, encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3,
                                                    "strides": [(2, 2), (2, 2), (1, 1)]})


def test_oracle_step_pin():
    """The fp64 oracle reproduces its own committed outputs (guards against silent drift of the checker)."""
    g0 = np.load(os.path.join(GOLD, "step_cfg1_small.npz"))
    m = O.OracleMVAE((8, 8, 3), [4, 2], encoder={"filters": [8, 8], "kernel_size": [(3, 3), (3, 3)],
                                                 "strides": [(2, 2), (1, 1)]}, sample_std=0.5, dtype=torch.float64, seed=11)
    m.compile(0.01, 1.0, 0.1)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 8, 8, 3, generator=g, dtype=torch.float64) * 255
    eps = [torch.randn(4, z, generator=g, dtype=torch.float64) for z in m.z_dims]
    res, grads = m.loss_and_grads(x, eps)
    assert np.allclose(res["loss"].detach().numpy(), g0["loss"], rtol=1e-10)
    assert np.allclose(res["out"].detach().numpy(), g0["out"], rtol=1e-9, atol=1e-9)
    assert np.allclose([float(grads[n].norm()) for n in sorted(grads)], g0["gnorm"], rtol=1e-8, atol=1e-12)


def test_adagrad_clipnorm_hand_value():
    m = O.OracleMVAE((8, 8, 3), [4, 2], encoder={"filters": [8], "kernel_size": [(3, 3)], "strides": [(1, 1)]},
                     dtype=torch.float64)
    m.compile(0.5, clip_norm=1.0)
    name = "encoder_0_conv_base/bias"
    w0 = m.params[name].clone()
    g = torch.zeros_like(w0)
    g[0] = 3.0
    g[1] = 4.0                      # norm 5 -> clipped to (0.6, 0.8)
    m.apply_grads({name: g}, {})
    exp0 = w0[0] - 0.5 * 0.6 / (np.sqrt(0.1 + 0.36) + 1e-7)
    exp1 = w0[1] - 0.5 * 0.8 / (np.sqrt(0.1 + 0.64) + 1e-7)
    assert abs(float(m.params[name][0] - exp0)) < 1e-12 and abs(float(m.params[name][1] - exp1)) < 1e-12


# ---- compile() losses, schedule and train(): vectors produced by EXECUTING the reference's own closures ------------
def _loss_cases():
    g = np.load(os.path.join(GOLD, "compile_losses.npz"))
    for n in sorted({k.split("_")[0] for k in g.files}):
        yield n, {k[len(n) + 1:]: g[k] for k in g.files if k.startswith(n + "_")}


def test_losses_match_reference_closures():
    """oracle r_loss / r_loss_metric / kl_loss against multiscale_vae.py:453-495 run on the numpy shim."""
    ncase = 0
    for name, c in _loss_cases():
        H, W, C = c["y"].shape[1:]
        m = O.OracleMVAE.__new__(O.OracleMVAE)
        m.input_dims = (H, W, C)
        y, yh = torch.from_numpy(c["y"]), torch.from_numpy(c["yh"])
        mu, lv = torch.from_numpy(c["mu"]), torch.from_numpy(c["lv"])
        rf, kf = c["factors"]
        assert np.allclose(m.r_loss_metric(y, yh).numpy(), c["vae_r_loss"], rtol=1e-12, atol=0), name
        assert np.allclose(m.r_loss(y, yh).numpy(), c["vae_r_experimental_loss"], rtol=1e-12, atol=0), name
        assert np.allclose(O.OracleMVAE.kl_loss(mu, lv).numpy(), c["vae_kl_loss"], rtol=1e-12, atol=0), name
        tot = m.r_loss(y, yh) * rf + O.OracleMVAE.kl_loss(mu, lv) * kf
        assert np.allclose(tot.numpy(), c["vae_loss"], rtol=1e-12, atol=0), name
        ncase += 1
    assert ncase == 5


def test_step_decay_matches_reference_schedule():
    """oracle.step_decay and the product's schedule.step_decay_schedule against schedule.py:7-21 run on the shim."""
    from multiscale_variational_autoencoder_b200 import schedule
    rows = np.load(os.path.join(GOLD, "schedule.npz"))["rows"]
    assert len(rows) == 32
    for lr0, decay, step, epoch, want in rows:
        assert O.step_decay(lr0, decay, step, epoch) == want
        assert schedule.step_decay_schedule(lr0, decay, int(step))(epoch) == want


def test_train_fit_call_of_the_reference():
    """What the reference's train() hands to keras fit (recorded by running multiscale_vae.py:508-557 on a stub): the
    product's train() follows these facts -- x is its own target, shuffled mini-batches, every sample of an epoch is used
    (fit runs the short last batch), epochs counted from initial_epoch, viz callback on the first 16 images."""
    import json
    with open(os.path.join(GOLD, "train_fit_call.json")) as f:
        calls = json.load(f)
    for c in calls.values():
        assert c["x_is_target"] and c["kwargs"]["shuffle"] is True
        assert c["callbacks"][:2] == ["LearningRateScheduler", "SaveIntermediateResultsCallback"]
        assert c["viz_first_n"] == 16 and c["viz_every"] == 100
    assert calls["resume_ckpt"]["kwargs"]["initial_epoch"] == 2
    assert calls["resume_ckpt"]["checkpoint_files"] == ["weights/weights-{epoch:03d}-{loss:.2f}.h5", "weights/weights.h5"]
    assert calls["plain"]["checkpoint_files"] == []
