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
    ((64, 64, 3), 6, 16),
    ((32, 32, 3), 5, 24),
    ((32, 32, 3), 3, 37),
    ((16, 32, 3), 2, 18),
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


@pytest.mark.parametrize("fold", ["", "both"])
@pytest.mark.parametrize("dims,levels,B", CASES)
def test_fused_chain_matches_layer_by_layer(dims, levels, B, fold):
    from multiscale_variational_autoencoder_b200 import _lib
    lib = _lib.load()
    z = [8] * levels
    cfg = dict(input_dims=dims, z_dims=z, sample_std=0.5,
               encoder={"filters": [32, 32, 32], "kernel_size": [(3, 3)] * 3, "strides": [(2, 2), (1, 1), (1, 1)]})
    model, _, x, eps = S.make_pair(cfg, B, precision="tf32", seed=3)
    model.compile(0.01, 1.0, 0.1)
    eng = S.run_product(model, x, eps, graph=False)
    eng.fold_se = fold            # squeeze-excite gate folded into the launches (opt-in) or as its own kernels
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
    # Two TF32 runs of one model are not comparable bit for bit: an fp32 summation-order difference of 1e-7 in a GAP sum
    # moves a TF32 operand across a rounding boundary (1e-3 of that element), the squeeze-excite BatchNorm (batch
    # statistics, eps 1e-3) spreads it over the batch, and ReLU masks flip behind it -- the same spread two runs of the
    # layer-by-layer path show (tests/test_gpu_tc.py docstring).  The kernels are held to 2e-5 call by call above; here the
    # engine plumbing (which tensor feeds which launch, for every level and chain) is checked at the TF32 tolerances of
    # the whole-step oracle tests: forward tensors 1e-3, gradients in the relative L2 norm.
    bad = []
    for k in ref:
        fwd = k.rsplit(".", 1)[-1] in ("a", "u", "gate", "y") or k.startswith("y") or k == "scalars"
        if fwd:
            e, tol = S.relerr(got[k], ref[k]), 1e-3
        else:       # a flipped ReLU mask changes single entries by O(1): relative L2 norm, as in tests/test_gpu_tc.py
            e = float((got[k].double() - ref[k].double()).norm() / max(float(ref[k].double().norm()), 1e-30))
            tol = 5e-2
        if e > tol:
            bad.append((e, k))
    assert not bad, sorted(bad, reverse=True)[:20]
    num = den = 0.0
    for k in gref:
        num += float((ggot[k].double() - gref[k].double()).pow(2).sum())
        den += float(gref[k].double().pow(2).sum())
    assert (num / den) ** 0.5 <= 3e-2, (num / den) ** 0.5


# B, H, W: every tile geometry, batches that leave the last tile partly empty; B*H*W >= 512 so that the layer-by-layer
# TF32 calls take the tensor-core kernels too (below that they fall back to fp32 CUDA cores)
KERNEL_CASES = [(5, 16, 16), (3, 32, 32), (2, 64, 64), (9, 8, 8), (37, 4, 4), (130, 2, 2), (520, 1, 1), (3, 16, 32), (2, 8, 128)]


@pytest.mark.parametrize("B,H,W", KERNEL_CASES)
def test_fused_kernels_match_layer_kernels(B, H, W):
    """mvae_mbv3_fused_fwd / _bwd call by call against the per-layer C-ABI calls they replace, on the same random
    operands (no squeeze-excite in between, so nothing amplifies fp32 summation-order noise)."""
    import ctypes as C
    from multiscale_variational_autoencoder_b200 import _lib as L
    lib = L.load()
    L.require_b200(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + W)
    rnd = lambda *sh, scale=1.0: (torch.randn(*sh, generator=g) * scale).to(dev).contiguous()
    s = torch.cuda.current_stream().cuda_stream
    Cc = 32
    ck = lambda rc: L.check(rc, "call")
    P = lambda t: t.data_ptr()
    d11 = L.ConvDesc(B, H, W, Cc, 1, 1, 1, 1, Cc, 0, L.PREC_TF32)
    relerr = lambda a, b: float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))
    shape = (B, H, W, Cc)
    w0, w2, w2p = rnd(Cc, Cc, scale=0.2), rnd(Cc, Cc, scale=0.2), rnd(Cc, Cc, scale=0.2)
    b0, b2, bd = rnd(Cc, scale=0.1), rnd(Cc, scale=0.1), rnd(Cc, scale=0.1)
    wd = rnd(3, 3, Cc, 1, scale=0.3)

    # ------------------------------------------------------------------------------------------------- forward
    u_prev, x_prev = rnd(*shape).relu_(), rnd(*shape)
    gate = torch.rand(B, Cc, generator=g).to(dev)
    y_ref, a_ref, u_ref = torch.empty(shape, device=dev), torch.empty(shape, device=dev), torch.empty(shape, device=dev)
    gap_ref = torch.zeros(B, Cc, device=dev)
    ck(lib.mvae_conv2d_fwd(C.byref(d11), P(u_prev), P(w2), P(b2), P(gate), P(x_prev), 0, P(y_ref), s))
    ck(lib.mvae_conv2d_fwd(C.byref(d11), P(y_ref), P(w0), P(b0), 0, 0, 1, P(a_ref), s))
    ck(lib.mvae_dwconv3x3_fwd(P(a_ref), P(wd), P(bd), P(u_ref), P(gap_ref), B, H, W, Cc, s))
    y, a, u = torch.zeros(shape, device=dev), torch.zeros(shape, device=dev), torch.zeros(shape, device=dev)
    gap = torch.zeros(B, Cc, device=dev)
    fa = L.Mbv3FwdArgs(B, H, W, Cc)
    fa.u_prev, fa.x_prev, fa.gate_prev, fa.w2, fa.b2, fa.y = P(u_prev), P(x_prev), P(gate), P(w2), P(b2), P(y)
    fa.w0, fa.b0, fa.wd, fa.bd, fa.a, fa.u, fa.gap_sum = P(w0), P(b0), P(wd), P(bd), P(a), P(u), P(gap)
    ck(lib.mvae_mbv3_fused_fwd(C.byref(fa), s))
    torch.cuda.synchronize()
    for name, got, ref in (("y", y, y_ref), ("a", a, a_ref), ("u", u, u_ref), ("gap", gap, gap_ref)):
        assert relerr(got, ref) <= 1e-5, ("F2F1", name, relerr(got, ref))
    # first launch of a chain (no F2 half): reads x
    a.zero_(), u.zero_(), gap.zero_()
    fa = L.Mbv3FwdArgs(B, H, W, Cc)
    fa.x, fa.w0, fa.b0, fa.wd, fa.bd, fa.a, fa.u, fa.gap_sum = P(y_ref), P(w0), P(b0), P(wd), P(bd), P(a), P(u), P(gap)
    ck(lib.mvae_mbv3_fused_fwd(C.byref(fa), s))
    torch.cuda.synchronize()
    for name, got, ref in (("a", a, a_ref), ("u", u, u_ref), ("gap", gap, gap_ref)):
        assert relerr(got, ref) <= 1e-5, ("F1", name, relerr(got, ref))
    # last launch (no F1 half)
    y.zero_()
    fa = L.Mbv3FwdArgs(B, H, W, Cc)
    fa.u_prev, fa.x_prev, fa.gate_prev, fa.w2, fa.b2, fa.y = P(u_prev), P(x_prev), P(gate), P(w2), P(b2), P(y)
    ck(lib.mvae_mbv3_fused_fwd(C.byref(fa), s))
    torch.cuda.synchronize()
    assert relerr(y, y_ref) <= 1e-5, ("F2", relerr(y, y_ref))

    # ------------------------------------------------------------------------------------------------ backward
    dy, dgap, up = rnd(*shape), rnd(B, Cc, scale=0.05), rnd(*shape).relu_()
    dv, da_ref, dx_ref, dvp = (torch.empty(shape, device=dev) for _ in range(4))
    dwd_ref, dbd_ref, dg_ref = torch.zeros(3, 3, Cc, 1, device=dev), torch.zeros(Cc, device=dev), torch.zeros(B, Cc, device=dev)
    ck(lib.mvae_conv2d_dgrad(C.byref(d11), P(dy), P(w2), 0, 0, 0, 0, P(dv), s))
    ck(lib.mvae_dwconv3x3_bwd(P(a_ref), P(u_ref), P(dv), P(gate), P(dgap), P(wd), P(da_ref), P(dwd_ref), P(dbd_ref), B, H, W,
                              Cc, s))
    ck(lib.mvae_conv2d_dgrad(C.byref(d11), P(da_ref), P(w0), 0, P(dy), 0, 0, P(dx_ref), s))
    ck(lib.mvae_conv2d_dgrad(C.byref(d11), P(dx_ref), P(w2p), 0, 0, 0, 0, P(dvp), s))
    ck(lib.mvae_se_dgate_reduce(P(dvp), P(up), P(dg_ref), B, H * W, Cc, s))
    da, dx = torch.zeros(shape, device=dev), torch.zeros(shape, device=dev)
    dwd, dbd, dg = torch.zeros_like(dwd_ref), torch.zeros_like(dbd_ref), torch.zeros_like(dg_ref)
    ba = L.Mbv3BwdArgs(B, H, W, Cc)
    ba.dy, ba.u, ba.a, ba.gate, ba.dgap, ba.w2, ba.wd, ba.w0 = P(dy), P(u_ref), P(a_ref), P(gate), P(dgap), P(w2), P(wd), P(w0)
    ba.da, ba.dx, ba.dwd, ba.dbd = P(da), P(dx), P(dwd), P(dbd)
    ba.w2_prev, ba.u_prev, ba.dgate_prev = P(w2p), P(up), P(dg)
    ck(lib.mvae_mbv3_fused_bwd(C.byref(ba), s))
    torch.cuda.synchronize()
    # (dgate: the fused launch adds the residual dy inside the tensor-core accumulator, the layer kernels in fp32 registers;
    # a last-bit difference of dx moves a few TF32 operand roundings of the following product, and a 2x2 image sums only
    # four of them per entry)
    for name, got, ref, tol in (("da", da, da_ref, 2e-5), ("dx", dx, dx_ref, 2e-5), ("dwd", dwd, dwd_ref, 2e-5),
                                ("dbd", dbd, dbd_ref, 2e-5), ("dgate", dg, dg_ref, 2e-4)):
        assert relerr(got, ref) <= tol, ("B2B1", name, relerr(got, ref))
    # first launch of the backward chain (no B2 half): dgate of the landed gradient
    dg.zero_()
    ba = L.Mbv3BwdArgs(B, H, W, Cc)
    ba.dy, ba.w2_prev, ba.u_prev, ba.dgate_prev = P(dx_ref), P(w2p), P(up), P(dg)
    ck(lib.mvae_mbv3_fused_bwd(C.byref(ba), s))
    torch.cuda.synchronize()
    assert relerr(dg, dg_ref) <= 2e-5, ("B1", relerr(dg, dg_ref))
    # last launch (no B1 half)
    da.zero_(), dx.zero_(), dwd.zero_(), dbd.zero_()
    ba = L.Mbv3BwdArgs(B, H, W, Cc)
    ba.dy, ba.u, ba.a, ba.gate, ba.dgap, ba.w2, ba.wd, ba.w0 = P(dy), P(u_ref), P(a_ref), P(gate), P(dgap), P(w2), P(wd), P(w0)
    ba.da, ba.dx, ba.dwd, ba.dbd = P(da), P(dx), P(dwd), P(dbd)
    ck(lib.mvae_mbv3_fused_bwd(C.byref(ba), s))
    torch.cuda.synchronize()
    for name, got, ref in (("da", da, da_ref), ("dx", dx, dx_ref), ("dwd", dwd, dwd_ref), ("dbd", dbd, dbd_ref)):
        assert relerr(got, ref) <= 2e-5, ("B2", name, relerr(got, ref))


@pytest.mark.parametrize("B,H,W", [(64, 16, 16), (41, 8, 8), (130, 4, 4), (300, 2, 2), (520, 1, 1), (70, 8, 16)])
def test_folded_gate_matches_se_kernels(B, H, W):
    """One mobilenetV3 block with the squeeze-excite gate folded into the tile launches (F1 -> F2, B1 -> B2) against the
    per-layer calls with mvae_se_gate_fwd / _bwd in between: activations, gate, moving statistics, data gradients,
    depthwise gradients, and the squeeze-excite weight gradients mvae_se_gate_bwd computes from the folded launches'
    scratch."""
    import ctypes as C
    from multiscale_variational_autoencoder_b200 import _lib as L
    lib = L.load()
    L.require_b200(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(B + 7 * H + W)
    rnd = lambda *sh, scale=1.0: (torch.randn(*sh, generator=g) * scale).to(dev).contiguous()
    s = torch.cuda.current_stream().cuda_stream
    Cc, HW = 32, H * W
    ck = lambda rc: L.check(rc, "call")
    P = lambda t: t.data_ptr()
    Z = lambda *sh: torch.zeros(*sh, device=dev)
    d11 = L.ConvDesc(B, H, W, Cc, 1, 1, 1, 1, Cc, 0, L.PREC_TF32)
    relerr = lambda a, b: float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))
    shape = (B, H, W, Cc)
    x = rnd(*shape)
    w0, w2 = rnd(Cc, Cc, scale=0.2), rnd(Cc, Cc, scale=0.2)
    b0, b2, bd = rnd(Cc, scale=0.1), rnd(Cc, scale=0.1), rnd(Cc, scale=0.1)
    wd = rnd(3, 3, Cc, 1, scale=0.3)
    s0, s1 = rnd(Cc, Cc, scale=0.4), rnd(Cc, Cc, scale=0.4)
    sb0, sb1 = rnd(Cc, scale=0.2), rnd(Cc, scale=0.2)
    gam, bet = 1.0 + rnd(Cc, scale=0.1), rnd(Cc, scale=0.1)
    nws = int(lib.mvae_se_gate_ws_floats(B, Cc))
    eps, mom = 1e-3, 0.99

    # ---- reference: layer by layer
    a_r, u_r, y_r, gate_r, gap_r = Z(*shape), Z(*shape), Z(*shape), Z(B, Cc), Z(B, Cc)
    mm_r, mv_r, ws_r = Z(Cc), torch.ones(Cc, device=dev), Z(nws)
    ck(lib.mvae_conv2d_fwd(C.byref(d11), P(x), P(w0), P(b0), 0, 0, 1, P(a_r), s))
    ck(lib.mvae_dwconv3x3_fwd(P(a_r), P(wd), P(bd), P(u_r), P(gap_r), B, H, W, Cc, s))
    ck(lib.mvae_se_gate_fwd(P(gap_r), P(s0), P(sb0), P(gam), P(bet), P(s1), P(sb1), P(mm_r), P(mv_r), P(gate_r), P(ws_r), B, Cc,
                            HW, eps, mom, 1, s))
    ck(lib.mvae_conv2d_fwd(C.byref(d11), P(u_r), P(w2), P(b2), P(gate_r), P(x), 0, P(y_r), s))
    dy = rnd(*shape)
    dv, da_r, dx_r, dg_r, dgap_r = Z(*shape), Z(*shape), Z(*shape), Z(B, Cc), Z(B, Cc)
    dwd_r, dbd_r = Z(3, 3, Cc, 1), Z(Cc)
    ser = [Z(Cc, Cc), Z(Cc), Z(Cc), Z(Cc), Z(Cc, Cc), Z(Cc)]          # ds0, dsb0, dgamma, dbeta, ds1, dsb1
    ck(lib.mvae_conv2d_dgrad(C.byref(d11), P(dy), P(w2), 0, 0, 0, 0, P(dv), s))
    ck(lib.mvae_se_dgate_reduce(P(dv), P(u_r), P(dg_r), B, HW, Cc, s))
    ck(lib.mvae_se_gate_bwd(P(dg_r), P(s0), P(gam), P(bet), P(s1), P(ws_r), P(dgap_r), *[P(t) for t in ser], B, Cc, HW, s))
    ck(lib.mvae_dwconv3x3_bwd(P(a_r), P(u_r), P(dv), P(gate_r), P(dgap_r), P(wd), P(da_r), P(dwd_r), P(dbd_r), B, H, W, Cc, s))
    ck(lib.mvae_conv2d_dgrad(C.byref(d11), P(da_r), P(w0), 0, P(dy), 0, 0, P(dx_r), s))

    # ---- folded: F1, F2, B1, B2
    a, u, y, gate = Z(*shape), Z(*shape), Z(*shape), Z(B, Cc)
    mm, mv, ws = Z(Cc), torch.ones(Cc, device=dev), Z(nws)
    stat = torch.zeros(64, dtype=torch.float64, device=dev)
    bstat = torch.zeros(64, dtype=torch.float64, device=dev)
    fa = L.Mbv3FwdArgs(B, H, W, Cc)
    fa.x, fa.w0, fa.b0, fa.wd, fa.bd, fa.a, fa.u = P(x), P(w0), P(b0), P(wd), P(bd), P(a), P(u)
    fa.se_w0, fa.se_b0, fa.se_ws, fa.se_stat = P(s0), P(sb0), P(ws), P(stat)
    ck(lib.mvae_mbv3_fused_fwd(C.byref(fa), s))
    fa = L.Mbv3FwdArgs(B, H, W, Cc)
    fa.u_prev, fa.x_prev, fa.w2, fa.b2, fa.y = P(u), P(x), P(w2), P(b2), P(y)
    fa.se_gamma_prev, fa.se_beta_prev, fa.se_w1_prev, fa.se_b1_prev = P(gam), P(bet), P(s1), P(sb1)
    fa.se_mm_prev, fa.se_mv_prev, fa.se_ws_prev, fa.gate_out_prev = P(mm), P(mv), P(ws), P(gate)
    fa.se_stat_prev = P(stat)
    fa.bn_eps, fa.bn_momentum, fa.training = eps, mom, 1
    ck(lib.mvae_mbv3_fused_fwd(C.byref(fa), s))
    torch.cuda.synchronize()
    for name, got, ref, tol in (("a", a, a_r, 1e-5), ("u", u, u_r, 1e-5), ("gate", gate, gate_r, 1e-4), ("y", y, y_r, 5e-4),
                                ("moving_mean", mm, mm_r, 1e-5), ("moving_var", mv, mv_r, 1e-5),
                                ("ws", ws[:2 * B * Cc], ws_r[:2 * B * Cc], 1e-5),
                                ("ws.s", ws[3 * B * Cc:4 * B * Cc], ws_r[3 * B * Cc:4 * B * Cc], 1e-4),
                                ("ws.stats", ws[6 * B * Cc:6 * B * Cc + 2 * Cc], ws_r[6 * B * Cc:6 * B * Cc + 2 * Cc], 1e-5)):
        assert relerr(got, ref) <= tol, ("fwd", name, relerr(got, ref))
    # backward on the REFERENCE forward state (same masks on both sides)
    da, dx, dg = Z(*shape), Z(*shape), Z(B, Cc)
    dwd, dbd = Z(3, 3, Cc, 1), Z(Cc)
    ba = L.Mbv3BwdArgs(B, H, W, Cc)
    ba.dy, ba.w2_prev, ba.u_prev, ba.dgate_prev = P(dy), P(w2), P(u_r), P(dg)
    ba.se_w1_prev, ba.se_ws_prev, ba.se_bstat_prev = P(s1), P(ws_r), P(bstat)
    ck(lib.mvae_mbv3_fused_bwd(C.byref(ba), s))
    ba = L.Mbv3BwdArgs(B, H, W, Cc)
    ba.dy, ba.u, ba.a, ba.gate, ba.w2, ba.wd, ba.w0 = P(dy), P(u_r), P(a_r), P(gate_r), P(w2), P(wd), P(w0)
    ba.da, ba.dx, ba.dwd, ba.dbd = P(da), P(dx), P(dwd), P(dbd)
    ba.se_w0, ba.se_gamma, ba.se_ws, ba.se_bstat = P(s0), P(gam), P(ws_r), P(bstat)
    ck(lib.mvae_mbv3_fused_bwd(C.byref(ba), s))
    seg = [Z(Cc, Cc), Z(Cc), Z(Cc), Z(Cc), Z(Cc, Cc), Z(Cc)]
    dgap2 = Z(B, Cc)
    ck(lib.mvae_se_gate_bwd(P(dg), P(s0), P(gam), P(bet), P(s1), P(ws_r), P(dgap2), *[P(t) for t in seg], B, Cc, HW, s))
    torch.cuda.synchronize()
    for name, got, ref in (("dgate", dg, dg_r), ("da", da, da_r), ("dx", dx, dx_r), ("dwd", dwd, dwd_r), ("dbd", dbd, dbd_r)):
        assert relerr(got, ref) <= 2e-4, ("bwd", name, relerr(got, ref))
    for i, (got, ref) in enumerate(zip(seg, ser)):
        assert relerr(got, ref) <= 2e-4, ("se weight gradient", i, relerr(got, ref))


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
