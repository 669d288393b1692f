"""Fused mobilenetV3 tile kernels (csrc/mbv3_fused.cu) against the layer-by-layer TF32 kernels they replace, on the same
engine, weights and inputs.  Both paths round the same operands to TF32 and accumulate in fp32, so they agree to fp32
summation-order noise (the GAP / gate-gradient sums are added in a different order); the layer-by-layer path itself is
held to the fp64 oracle in tests/test_gpu_tc.py, and the whole-step oracle tests there run through the fused path."""
import pytest
import torch

import test_gpu_step as S

pytestmark = pytest.mark.gpu

# every tile geometry: strips with halo (64x64: 4 rows + 2, 32x32: 8 rows + 2), whole images (16x16 ... 1x1), and batches
# that leave the last tile partly empty
CASES = [
    # input, levels, B
    ((64, 64, 3), 6, 3),
    ((32, 32, 3), 5, 6),
    ((32, 32, 3), 3, 37),
    ((16, 32, 3), 2, 5),
]


def _snapshot(eng):
    out = {}
    for side, lists in (("enc", eng.enc_ops), ("dec", eng.dec_ops)):
        for i, ops in enumerate(lists):
            for k, op in enumerate(ops):
                for j, m in enumerate(getattr(op, "blocks", [])):
                    tag = f"{side}{i}.{k}.{j}"
                    out[tag + ".a"] = m.a.clone()
                    out[tag + ".u"] = m.u.clone()
                    out[tag + ".gate"] = m.gate.clone()
                    out[tag + ".y"] = m.y.data.clone()
                    out[tag + ".da"] = m.da.clone()
                    out[tag + ".dx"] = m.x.grad.clone()
                    out[tag + ".dgap"] = m.dgap.clone()
    for i, t in enumerate(eng.ys):
        out[f"y{i}"] = t.data.clone()
    out["scalars"] = eng.scalars.clone()
    return out


@pytest.mark.parametrize("dims,levels,B", CASES)
def test_fused_chain_matches_layer_by_layer(dims, levels, B):
    from multiscale_variational_autoencoder_b200 import _lib
    lib = _lib.load()
    z = [8] * levels
    cfg = dict(input_dims=dims, z_dims=z, sample_std=0.5,
               encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (1, 1), (1, 1)]})
    model, _, x, eps = S.make_pair(cfg, B, precision="tf32", seed=3)
    model.compile(0.01, 1.0, 0.1)
    eng = S.run_product(model, x, eps, graph=False)
    chains = [op for ops in eng.enc_ops + eng.dec_ops for op in ops if hasattr(op, "blocks")]
    assert chains and all(c.fused() for c in chains), "no fused chain was built"
    assert sum(len(c.blocks) for c in chains) == 6 * levels

    def run(fuse):
        eng.fuse_mbv3 = fuse
        before = lib.mvae_tc_launch_count()
        eng.forward_train()
        eng.backward()
        torch.cuda.synchronize()
        return _snapshot(eng), {k: v.clone() for k, v in model._ps.state_dict(grads=True).items()}, \
            lib.mvae_tc_launch_count() - before

    ref, gref, n_ref = run(False)
    got, ggot, n_got = run(True)
    # every block went through the fused kernels: (n + 1) launches forward and backward per chain, on top of the rest
    assert n_got >= 2 * sum(len(c.blocks) + 1 for c in chains), (n_got, n_ref)
    for k in ref:
        assert S.relerr(got[k], ref[k]) <= 2e-5, (k, S.relerr(got[k], ref[k]))
    for k in gref:
        scale = max(float(gref[k].abs().max()), 1e-6)
        err = float((ggot[k] - gref[k]).abs().max())
        assert err <= 5e-5 * scale + 1e-7, (k, err, scale)


def test_fused_entry_points_reject_unsupported_shapes():
    from multiscale_variational_autoencoder_b200 import _lib
    lib = _lib.load()
    assert lib.mvae_mbv3_fused_supported(8, 16, 16, 32, 32) == 1
    assert lib.mvae_mbv3_fused_supported(8, 32, 32, 32, 32) == 1
    assert lib.mvae_mbv3_fused_supported(8, 1, 1, 32, 32) == 1
    assert lib.mvae_mbv3_fused_supported(8, 16, 16, 64, 64) == 0      # wide filters: layer-by-layer kernels
    assert lib.mvae_mbv3_fused_supported(8, 12, 12, 32, 32) == 0      # 144 pixels do not tile 256
    assert lib.mvae_mbv3_fused_supported(8, 256, 256, 32, 32) == 0    # rows wider than 128 pixels
    a = _lib.Mbv3FwdArgs(8, 12, 12, 32)
    assert lib.mvae_mbv3_fused_fwd(a, None) == -3
