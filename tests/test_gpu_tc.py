"""tcgen05 (MVAE_PREC_TF32) parity.

Forward quantities (per-scale reconstructions, latents, ELBO terms) are held to 1e-3 against the fp64 oracle, as the
north star asks.  Gradients need care: a TF32 forward flips the ReLU / hard_sigmoid mask of every pre-activation that
lies within ~3e-4 of the kink (about 2.4e-4 of all elements for unit-variance activations), and each flip changes one
activation-gradient entry by O(1) -- an L2 error of sqrt(2.4e-4) ~ 2 % that no TF32 implementation can avoid and that
has nothing to do with kernel correctness (measured: scripts/debug_mbv3.py).  The gradient kernels
are therefore checked where the question is well posed: TF32 backward against FP32 backward FROM THE SAME SAVED
ACTIVATIONS (identical masks), tolerance 2e-3 per tensor; the end-to-end TF32-vs-oracle gradient error is bounded in
the relative L2 norm.  precision="fp32" meets 1e-4 on every gradient unconditionally (tests/test_gpu_step.py)."""
import pytest
import torch

import test_gpu_kernels as K
import test_gpu_step as S
from oracle import mvae_oracle as O

pytestmark = pytest.mark.gpu

TOL_TF32 = 1e-3

TC_CASES = [
    # B, H, W, Cin, k, s, Cout, act, coord, gate, residual
    (4, 32, 32, 32, 3, 2, 32, 0, 0, False, False),    # strided encoder conv / Conv2DTranspose pair
    (8, 16, 16, 32, 1, 1, 32, 1, 0, False, False),    # mbv3 conv0 + ReLU
    (8, 16, 16, 32, 1, 1, 32, 0, 0, True, True),      # mbv3 conv2: gate + residual
    (4, 16, 16, 64, 3, 1, 128, 0, 0, False, False),   # wide, several channel groups
    (2, 24, 20, 32, 3, 2, 64, 0, 0, False, False),    # ragged tile tail
    (3, 16, 16, 96, 5, 2, 32, 2, 0, False, False),    # 5x5, ELU
    (2, 16, 16, 64, 3, 1, 256, 0, 0, False, False),   # N = 256 (wgrad only: fwd/dgrad take N <= 128)
    (256, 1, 1, 2048, 1, 1, 256, 0, 0, False, False), # Dense heads: wgrad on tensor cores, fwd/dgrad split-K fp32
    # TMA tile geometries: tb images x th rows x tw pixels = 128 GEMM rows
    (32, 4, 4, 32, 3, 1, 32, 1, 0, False, False),     # 8 whole images per tile
    (36, 8, 8, 32, 3, 2, 32, 0, 0, False, False),     # strided, 4x4 outputs, last tile runs past the batch
    (9, 8, 8, 32, 3, 1, 64, 0, 0, False, True),       # 2 images per tile, 4.5 tiles, residual
    (2, 4, 128, 32, 3, 1, 32, 2, 0, False, False),    # one 128-pixel row segment per tile
    (3, 16, 256, 32, 1, 1, 32, 1, 0, True, True),     # flat 1x1 rows, gate + residual
    (5, 10, 13, 64, 1, 1, 32, 0, 0, True, False),     # flat 1x1, M = 650: partial last tile, gate across image boundaries
]


def _lib():
    from multiscale_variational_autoencoder_b200 import _lib
    _lib.require_b200(0)
    return _lib.load()


@pytest.mark.parametrize("case", TC_CASES)
def test_conv2d_tc(case):
    lib = _lib()
    before = lib.mvae_tc_launch_count()
    K.test_conv2d_fwd_dgrad_wgrad(lib, case, prec=1, tol=TOL_TF32)
    assert lib.mvae_tc_launch_count() >= before + 1, "tensor-core path was not taken"


def test_conv2d_transpose_tc():
    lib = _lib()
    before = lib.mvae_tc_launch_count()
    K.test_conv2d_transpose_layer(lib, 8, 8, 8, 32, 3, 2, 32, prec=1, tol=TOL_TF32)
    assert lib.mvae_tc_launch_count() >= before + 2


def _grads(ps):
    return {k: v.clone() for k, v in ps.state_dict(grads=True).items()}


def _cmp_grads(a, b, tol, what, skip=()):
    gmax = max(float(v.abs().max()) for v in b.values())
    bad = []
    for k in b:
        if k in skip:
            continue
        e = float((a[k] - b[k]).abs().max() / max(float(b[k].abs().max()), 1e-3 * gmax))
        if e > tol:
            bad.append((k, e))
    assert not bad, (what, bad[:10])


@pytest.mark.parametrize("B,H,W,Cc,F", [(8, 16, 16, 32, 32), (4, 16, 16, 64, 128)])
def test_mobilenetv3_block_tc(B, H, W, Cc, F):
    from multiscale_variational_autoencoder_b200 import engine as E
    lib = _lib()
    before = lib.mvae_tc_launch_count()
    ps = E.ParamStore(torch.device("cuda", 0), seed=3)
    E.declare_mbv3(ps, "m_", Cc, F)
    ps.finalize()
    sd = {k: v.double() for k, v in ps.state_dict().items()}
    m = O.OracleMVAE.__new__(O.OracleMVAE)
    m.params = sd
    x = K.rnd((B, H, W, Cc), 1)
    y = m._mbv3(x.double(), "m_", True, {})
    eng = K.MiniTrain(ps, B, prec=1)
    xt = E.T(x.cuda(), torch.empty((B, H, W, Cc), device="cuda"))
    op = E.MobileNetV3(eng, xt, "m_", F)
    op.fwd()
    assert K.relerr(op.y.data, y) <= TOL_TF32, "fwd vs oracle"
    op.y.grad.copy_(K.rnd((B, H, W, Cc), 2).cuda())
    op.bwd()
    g_tf32, dx_tf32 = _grads(ps), xt.grad.clone()
    assert lib.mvae_tc_launch_count() >= before + 4
    # same saved activations, FP32 kernels
    op.d0.precision = op.d2.precision = 0
    ps.grads.zero_()
    for t in eng._keep:
        t.zero_()
    op.gap  # (forward accumulators are not needed by the backward)
    op.bwd()
    _cmp_grads(g_tf32, _grads(ps), 2e-3, "mbv3 parameter gradients, TF32 vs FP32 backward")
    assert K.relerr(dx_tf32, xt.grad) <= 2e-3, "dx"


@pytest.mark.parametrize("name,B", [("cfg1", 16), ("cfg2", 16)])
def test_step_parity_tf32(name, B):
    lib = _lib()
    before = lib.mvae_tc_launch_count()
    cfg = S.CFGS[name]
    model, oracle, x, eps = S.make_pair(cfg, B, precision="tf32")
    model.compile(0.01, 1.0, 0.1)
    oracle.compile(0.01, 1.0, 0.1)
    eng = S.run_product(model, x, eps, graph=False)
    eng.forward_train()
    eng.backward()
    torch.cuda.synchronize()
    assert lib.mvae_tc_launch_count() > before + 10
    res, grads = oracle.loss_and_grads(x.double(), [e.double() for e in eps])
    for i, z in enumerate(cfg["z_dims"]):
        mulv = eng.mulv[i].data.view(B, 2 * z)
        assert S.relerr(mulv[:, :z], res["mu"][i]) <= TOL_TF32, ("mu", i)
        assert S.relerr(mulv[:, z:], res["log_var"][i]) <= TOL_TF32, ("log_var", i)
        assert S.relerr(eng.zT[i].data.view(B, z), res["z"][i]) <= TOL_TF32, ("z", i)
        assert S.relerr(eng.ys[i].data, res["y"][i]) <= TOL_TF32, ("y", i)
        assert S.relerr(eng.kl[i], res["kl_per_scale"][i]) <= TOL_TF32, ("kl", i)
    # merged output in raw [0,255] units: the per-level errors (each <= 1e-3) add up -> 2.5e-3 stated for this tensor
    assert S.relerr(eng.out, res["out"]) <= 2.5e-3
    assert S.relerr(eng.per_sample[0], res["r_loss"]) <= TOL_TF32
    assert S.relerr(eng.per_sample[2], res["kl_loss"]) <= TOL_TF32
    g_tf32 = _grads(model._ps)
    # (1) kernels: TF32 backward == FP32 backward on the same saved activations
    eng.set_precision(0)
    eng.backward()
    torch.cuda.synchronize()
    # (biases feeding the decoder BatchNorm have an exactly-zero gradient: only summation noise, see test_gpu_step.py)
    nlast = len(cfg["encoder"]["filters"]) - 1
    skip = {f"decoder_{i}__{nlast}_mobilenetV3_conv2/bias" for i in range(len(cfg["z_dims"]))}
    _cmp_grads(g_tf32, _grads(model._ps), 2e-3, "TF32 vs FP32 backward, same activations", skip)
    # (2) end to end against the oracle: bounded by the ReLU mask flips of the TF32 forward (see module docstring)
    num = den = 0.0
    for k, g in grads.items():
        w = oracle.params[k].detach()
        if oracle.reg[k] == O.REG_L1:
            g = g - O.REG_FACTOR * torch.sign(w)
        elif oracle.reg[k] == O.REG_L2:
            g = g - 2 * O.REG_FACTOR * w
        num += float((g_tf32[k].double() - g).pow(2).sum())
        den += float(g.pow(2).sum())
    assert (num / den) ** 0.5 <= 5e-2, (num / den) ** 0.5
