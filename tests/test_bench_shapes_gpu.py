"""Parity at the shapes and kernel instantiations bench.py actually times
("What's weak" 1-3): AR-FPN gather / apply at C = 256 on the
full Faster R-CNN pyramid (NV = 2 fp32, V = 8 bf16) and on BASELINE config 2's
RetinaNet pyramid (B = 8), the whole workload.TrainStep against the oracle's
composition of the same step, and the committed golden fixtures (minted from
the compiled reference) through the CUDA path."""
import os

import numpy as np
import pytest
import torch

from util import STRIDES, assert_close_bf16, assert_close_fp32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _sum_tol(got, ref, what, rel=1e-5, scale=2e-5):
    """Gradients that are sums of many terms taken in another order than the
    reference's (its backward is an atomicAdd scatter, ours a fixed-order pull;
    the channel reductions of the gate maps likewise): 1e-5 relative plus
    2e-5 of the tensor's largest magnitude."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = rel * ref.abs() + scale * float(ref.abs().max()) + 1e-6
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {float(err.max()):.3e}, " \
                          f"max ref {float(ref.abs().max()):.3e}"


# ---------------------------------------------------------------- AR-FPN, full size
CONFIGS = {
    # BASELINE configs[1]: Faster R-CNN R50, 2 img/GPU, 800x1344, strides 4..64
    "frcnn_b2": dict(batch=2, strides=(4, 8, 16, 32, 64)),
    # BASELINE configs[2]: RetinaNet R50 + AR-FPN, batch 8, strides 8..128
    # (configs/_base_/models/retinanet_r50_drfpn.py:14-25: FPN start_level=1, 5 outs)
    "retina_b8": dict(batch=8, strides=(8, 16, 32, 64, 128)),
}


def _fpn_inputs(oracle, cfg, seed=0):
    shapes = oracle.pyramid_shapes(800, 1344, cfg["strides"])
    B, C = cfg["batch"], 256
    xs = oracle.synthetic_pyramid(B, C, shapes, seed=seed)
    gen = torch.Generator().manual_seed(seed + 100)
    hr, wr = shapes[2]
    bsf = torch.randn(B, C, hr, wr, generator=gen)
    g1 = [torch.randn(B, 1, h, w, generator=gen) for h, w in shapes]
    g2 = [torch.randn(B, 1, h, w, generator=gen) for h, w in shapes]
    return shapes, xs, bsf, g1, g2


@pytest.mark.parametrize("cfg", list(CONFIGS))
@pytest.mark.parametrize("mode", ["f32_nhwc", "f32_nchw", "bf16_nhwc"])
def test_arfpn_at_bench_shapes(oracle, cuda, cfg, mode):
    import arfe_b200 as A
    shapes, xs, bsf, g1, g2 = _fpn_inputs(oracle, CONFIGS[cfg], seed=11)
    bf16 = mode.startswith("bf16")
    nhwc = mode.endswith("nhwc")
    if bf16:
        rnd = lambda t: t.bfloat16().float()
        xs, bsf, g1, g2 = [rnd(t) for t in xs], rnd(bsf), [rnd(t) for t in g1], [rnd(t) for t in g2]
    req = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    xo, g1o, g2o, bo = req(xs), req(g1), req(g2), bsf.clone().requires_grad_(True)
    ref_g = oracle.wfpn_gather(xo, 2)
    ref_o = oracle.wfpn_apply(xo, bo, g1o, g2o)
    gen = torch.Generator().manual_seed(12)
    gg = torch.randn(ref_g.shape, generator=gen)
    gs = [torch.randn(r.shape, generator=gen) for r in ref_o]
    if bf16:
        gg, gs = gg.bfloat16().float(), [t.bfloat16().float() for t in gs]
    # d x through the gather and through the residual separately (autograd would add them)
    gx_gather = torch.autograd.grad(ref_g, xo, gg, retain_graph=True, allow_unused=True)
    torch.autograd.backward(list(ref_o), gs)

    dt = torch.bfloat16 if bf16 else torch.float32
    mv = (lambda t: _cl(t.to(cuda, dt))) if nhwc else (lambda t: t.to(cuda, dt).contiguous())
    xg = [mv(x).requires_grad_(True) for x in xs]
    got_g = A.fpn_gather(xg, 2)
    gxg = torch.autograd.grad(got_g, xg, mv(gg))
    bg = mv(bsf).requires_grad_(True)
    g1g = [t.to(cuda, dt).requires_grad_(True) for t in g1]
    g2g = [t.to(cuda, dt).requires_grad_(True) for t in g2]
    got_o = A.fpn_apply(xg, bg, g1g, g2g)
    torch.autograd.backward(list(got_o), [mv(g) for g in gs])
    torch.cuda.synchronize()
    if not bf16:
        # same op order as the reference (adds in level order, true division): exact
        assert torch.equal(got_g.cpu(), ref_g.detach()), float((got_g.cpu() - ref_g).abs().max())
    else:
        assert_close_bf16(got_g, ref_g, "gather bf16")
    for l in range(5):
        close = assert_close_bf16 if bf16 else assert_close_fp32
        close(got_o[l], ref_o[l], f"{cfg} {mode} apply out level {l}")
        rg = gx_gather[l] if gx_gather[l] is not None else torch.zeros_like(xs[l])
        close(gxg[l], rg, f"{cfg} {mode} gather dx level {l}")
        # d x of the residual is d out itself
        close(xg[l].grad, gs[l], f"{cfg} {mode} apply dx level {l}")
        for name, a, b in (("dg1", g1g, g1o), ("dg2", g2g, g2o)):
            _sum_tol(a[l].grad, b[l].grad, f"{cfg} {mode} {name} level {l}",
                     rel=1e-2 if bf16 else 1e-5, scale=1e-2 if bf16 else 2e-5)
    _sum_tol(bg.grad, bo.grad, f"{cfg} {mode} dbsf", rel=1e-2 if bf16 else 1e-5,
             scale=1e-2 if bf16 else 2e-5)


# ------------------------------------------------ the whole bench step vs the oracle's composition
def _step_host(wl, channels_last, order, seed=3):
    return wl.host_inputs(batch=2, rois_per_img=32, channels=256, img_h=256, img_w=320, seed=seed,
                          channels_last=channels_last, roi_order=order, smin=8.0, smax=230.0)


@pytest.mark.parametrize("variant", ["cl_split", "cl_cat", "nchw", "cl_split_interleaved"])
def test_train_step_matches_reference_step(oracle, cuda, variant):
    """workload.TrainStep.step() -- the thing bench.py times, C = 256 instantiations,
    plan and tile bins built on the second stream (plan_ready 1 / 2) -- against
    oracle.reference_step on the same host inputs."""
    from arfe_b200 import workload as wl
    cl = variant.startswith("cl")
    order = "interleaved" if variant.endswith("interleaved") else "image_major"
    rlev = 4
    host = _step_host(wl, cl, order)
    st = wl.TrainStep(host, cuda, channels_last=cl, split=(variant.startswith("cl_split")), roi_levels=rlev)
    assert st.overlap_plan == cl
    st.step()
    st.step()
    torch.cuda.synchronize()
    nchw = lambda v: [nchw(t) for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else \
        (v.contiguous() if torch.is_tensor(v) else v)
    ref = oracle.reference_step({k: nchw(v) for k, v in host.items()}, roi_levels=rlev)
    C = st.C
    assert torch.equal(st.gathered.cpu(), ref["gathered"]), "gather not exact"
    for l in range(5):
        assert_close_fp32(st.y[l], ref["y"][l], f"y level {l}")
    F = torch.cat([t for t in st.Fr], 1) if st.split else st.F
    assert_close_fp32(F, ref["F"], "RoI features (ori | lw | lh)")
    assert_close_fp32(st.z, ref["z"], "gated RoI features")
    d_ori = st.d_ori if st.split else st.dF[:, :C]
    assert_close_fp32(d_ori, ref["d_ori"], "d ori")
    assert_close_fp32(st.d_ab, ref["d_ab"], "d (a + b)")
    for l in range(5):
        _sum_tol(st.dy[l], ref["dy"][l], f"d y level {l}")
        _sum_tol(st.dx[l], ref["dx"][l], f"d x level {l}")
        _sum_tol(st.dg1[l], ref["dg1"][l], f"d g1 level {l}")
        _sum_tol(st.dg2[l], ref["dg2"][l], f"d g2 level {l}")
    _sum_tol(st.dbsf, ref["dbsf"], "d bsf")


# ------------------------------------------------------------ golden fixtures through CUDA
@pytest.mark.parametrize("mode", ["nchw", "nhwc", "split"])
def test_golden_arrff_through_cuda(cuda, mode):
    """tests/golden/arrff_small.npz (outputs of the reference's own RoIAlign
    compiled unmodified, make_golden.py) against the CUDA path: boxes and levels
    bit-exact, features and pyramid gradients within tolerance."""
    import arfe_b200 as A
    d = np.load(os.path.join(GOLD, "arrff_small.npz"))
    t = lambda k: torch.from_numpy(d[k]).to(cuda)
    rois = t("rois")
    mv = (lambda x: x) if mode == "nchw" else _cl
    feats = [mv(t(f"feat{l}")).requires_grad_(True) for l in range(5)]
    scales = [1.0 / s for s in STRIDES]
    dbg = A.roi_fuse_debug(rois, [f.shape[2] for f in feats], [f.shape[3] for f in feats], scales, 7, 0, 3)
    assert np.array_equal(dbg["boxes"].cpu().numpy(), d["boxes"])
    assert np.array_equal(dbg["lvl"].cpu().numpy(), d["lvls"])
    if mode == "split":
        out = torch.cat(A.roi_fuse_split(feats, rois, 7, scales, regions=3), 1)
    else:
        out = A.roi_fuse(feats, rois, 7, scales, regions=3, out_channels_last=(mode == "nhwc"))
    assert_close_fp32(out, torch.from_numpy(d["out"]), f"golden AR-RFF features ({mode})")
    out.backward(t("grad_out"))
    for l in range(5):
        g = feats[l].grad if feats[l].grad is not None else torch.zeros_like(feats[l])
        _sum_tol(g, torch.from_numpy(d[f"dfeat{l}"]), f"golden d feat{l} ({mode})")


@pytest.mark.parametrize("nhwc", [False, True])
def test_golden_roi_align_op_through_cuda(cuda, nhwc):
    """tests/golden/roi_align_op.npz: the operator twin of roi_align_ext.forward_v2 /
    backward_v2 (sample_num 0 and 2) on the reference's gradcheck sizes."""
    import arfe_b200 as A
    d = np.load(os.path.join(GOLD, "roi_align_op.npz"))
    rois = torch.from_numpy(d["rois"]).to(cuda)
    for sn in (0, 2):
        feat = torch.from_numpy(d["feat"]).to(cuda)
        feat = (_cl(feat) if nhwc else feat).requires_grad_(True)
        out = A.roi_align(feat, rois, 3, 1 / 8, sn, True)
        assert_close_fp32(out, torch.from_numpy(d[f"out_sn{sn}"]), f"golden roi_align sn={sn}")
        out.backward(torch.from_numpy(d[f"gout_sn{sn}"]).to(cuda))
        _sum_tol(feat.grad, torch.from_numpy(d[f"gin_sn{sn}"]), f"golden roi_align grad sn={sn}")
        layer = A.RoIAlign(3, 1 / 8, sample_num=sn)
        assert torch.equal(layer(feat.detach(), rois), out.detach())


def test_golden_level_thresholds_through_cuda(cuda):
    import arfe_b200 as A
    d = np.load(os.path.join(GOLD, "level_thresholds.npz"))
    rois = torch.from_numpy(d["rois"]).to(cuda)
    dbg = A.roi_fuse_debug(rois, [200, 100, 50, 25, 13], [336, 168, 84, 42, 21],
                           [1.0 / s for s in STRIDES], 7, 0, 1)
    assert np.array_equal(dbg["lvl"][0].cpu().numpy(), d["lvls"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_fpn_backward_equals_the_two_call_path(cuda, dtype):
    """arfe_fpn_backward_fused == arfe_fpn_apply_backward + arfe_fpn_gather_backward_acc(addend = d out),
    bit for bit (same arithmetic, same order), on the bench pyramid with a ragged top level."""
    from arfe_b200 import _lib as L
    from arfe_b200 import workload as wl
    host = wl.host_inputs(batch=2, rois_per_img=16, channels=256, img_h=416, img_w=672, dtype=dtype,
                          channels_last=True, seed=4)   # 104x168 ... 7x11: ratios 4, 2, 1, ~2, ~3.7
    st = wl.TrainStep(host, cuda)
    st.step()
    torch.cuda.synchronize()
    fused = [t.clone() for t in st.dx + st.dg1 + st.dg2 + [st.dbsf]]
    for t in st.dx + st.dg1 + st.dg2 + [st.dbsf]:
        t.fill_(float("nan"))
    st.glue_before_apply_bwd()
    L.check(st.fpn_apply_bwd(), "apply bwd")
    L.check(st.fpn_gather_bwd(), "gather bwd acc")
    torch.cuda.synchronize()
    two = st.dx + st.dg1 + st.dg2 + [st.dbsf]
    n = len(st.dx)
    for i, (a, b) in enumerate(zip(fused, two)):
        if i < n or i == len(fused) - 1:
            # d x and d bsf: same additions in the same order (for bf16 the two-call path reads d out
            # rounded to bf16 by the glue copy, the fused one reads the fp32 accumulators: d bsf differs)
            if dtype == torch.bfloat16 and i >= n:
                assert torch.allclose(a.float(), b.float(), rtol=2e-2, atol=2e-2 * float(b.abs().max())), i
            else:
                assert torch.equal(a, b), (i, float((a.float() - b.float()).abs().max()))
        else:
            # d gate maps: channel sums taken in another order (per-footprint reduction vs shuffle tree)
            tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
            assert float((a - b).abs().max()) <= tol * float(b.abs().max()) + 1e-6, i


@pytest.mark.parametrize("case", ["small_f32", "bench_f32", "bench_bf16"])
def test_fused_gate_convs_match_conv2d(cuda, case):
    """arfe_fpn_gate_conv_forward (both C -> 1 3x3 gate convolutions of every
    level, one pass over x) against torch CPU conv2d -- the op the reference's ConvModule runs."""
    import torch.nn.functional as F
    import arfe_b200 as A
    from arfe_b200 import workload as wl
    bf16 = case.endswith("bf16")
    if case.startswith("small"):
        B, C, shapes = 2, 24, [(9, 13), (5, 7), (3, 4), (2, 2), (1, 1)]
    else:
        B, C, shapes = 2, 256, wl.pyramid_shapes(800, 1344)
    gen = torch.Generator().manual_seed(8)
    xs = [torch.randn(B, C, h, w, generator=gen) for h, w in shapes]
    w1 = [torch.randn(1, C, 3, 3, generator=gen) * 0.05 for _ in shapes]
    w2 = [torch.randn(1, C, 3, 3, generator=gen) * 0.05 for _ in shapes]
    b1 = [torch.randn(1, generator=gen) for _ in shapes]
    b2 = [torch.randn(1, generator=gen) for _ in shapes]
    if bf16:
        xs = [x.bfloat16().float() for x in xs]
    dt = torch.bfloat16 if bf16 else torch.float32
    g1, g2 = A.fpn_gate_conv([_cl(x.to(cuda, dt)) for x in xs], [t.to(cuda) for t in w1], [t.to(cuda) for t in b1],
                             [t.to(cuda) for t in w2], [t.to(cuda) for t in b2])
    for l in range(5):
        for got, w, b in ((g1[l], w1[l], b1[l]), (g2[l], w2[l], b2[l])):
            ref = F.conv2d(xs[l], w, b, padding=1)
            assert got.shape == ref.shape and got.dtype == dt
            err = (got.float().cpu() - ref).abs().max()
            tol = (1e-2 if bf16 else 1e-5) * float(ref.abs().max()) + 1e-6   # 9 C-term sums in another order
            assert float(err) <= tol, (case, l, float(err), float(ref.abs().max()))


@pytest.mark.parametrize("case", ["small_f32", "odd_f32", "bench_f32", "bench_bf16"])
def test_fused_gate_convs_backward_matches_conv2d(cuda, case):
    """SURVEY 8(f) row 2, backward: arfe_fpn_gate_conv_backward (d x, d weights, d biases of both
    C -> 1 3x3 gate convolutions of every level in one pass over x) against torch CPU autograd of
    conv2d.  Sums over pixels (weights) and over 18 taps (d x) in another order than the
    library's: 2e-5 of the tensor's scale in fp32; bf16 I/O 1e-2."""
    import torch.nn.functional as F
    import arfe_b200 as A
    from arfe_b200 import workload as wl
    bf16 = case.endswith("bf16")
    if case.startswith("small"):
        B, C, shapes = 2, 24, [(9, 13), (5, 7), (3, 4), (2, 2), (1, 1)]
    elif case.startswith("odd"):
        B, C, shapes = 3, 132, [(7, 5), (4, 3), (1, 2)]          # C not a multiple of 128: ragged channel group
    else:
        B, C, shapes = 2, 256, wl.pyramid_shapes(800, 1344)
    n = len(shapes)
    gen = torch.Generator().manual_seed(9)
    xs = [torch.randn(B, C, h, w, generator=gen) for h, w in shapes]
    w1 = [torch.randn(1, C, 3, 3, generator=gen) * 0.05 for _ in shapes]
    w2 = [torch.randn(1, C, 3, 3, generator=gen) * 0.05 for _ in shapes]
    b1 = [torch.randn(1, generator=gen) for _ in shapes]
    b2 = [torch.randn(1, generator=gen) for _ in shapes]
    u1 = [torch.randn(B, 1, h, w, generator=gen) for h, w in shapes]
    u2 = [torch.randn(B, 1, h, w, generator=gen) for h, w in shapes]
    if bf16:
        xs = [x.bfloat16().float() for x in xs]
        u1 = [t.bfloat16().float() for t in u1]
        u2 = [t.bfloat16().float() for t in u2]
    dt = torch.bfloat16 if bf16 else torch.float32
    leaf = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    xr, w1r, w2r, b1r, b2r = leaf(xs), leaf(w1), leaf(w2), leaf(b1), leaf(b2)
    loss = sum((F.conv2d(xr[l], w1r[l], b1r[l], padding=1) * u1[l]).sum() +
               (F.conv2d(xr[l], w2r[l], b2r[l], padding=1) * u2[l]).sum() for l in range(n))
    loss.backward()
    xg = [_cl(x.to(cuda, dt)).requires_grad_(True) for x in xs]
    par = lambda ts: [t.to(cuda).requires_grad_(True) for t in ts]
    w1g, w2g, b1g, b2g = par(w1), par(w2), par(b1), par(b2)
    g1, g2 = A.fpn_gate_conv(xg, w1g, b1g, w2g, b2g)
    loss = sum((g1[l].float() * u1[l].to(cuda)).sum() + (g2[l].float() * u2[l].to(cuda)).sum() for l in range(n))
    loss.backward()
    tol = 1e-2 if bf16 else 2e-5
    for l in range(n):
        for name, got, want in (("dx", xg[l].grad, xr[l].grad), ("dw1", w1g[l].grad, w1r[l].grad),
                                ("dw2", w2g[l].grad, w2r[l].grad), ("db1", b1g[l].grad, b1r[l].grad),
                                ("db2", b2g[l].grad, b2r[l].grad)):
            assert got is not None and got.shape == want.shape, (name, l)
            err = float((got.float().cpu() - want).abs().max())
            assert err <= tol * float(want.abs().max()) + 1e-5, (case, name, l, err, float(want.abs().max()))
    # inputs that need no gradient: parameters only
    xn = [_cl(x.to(cuda, dt)) for x in xs]
    w1n = par(w1)
    g1, g2 = A.fpn_gate_conv(xn, w1n, b1g, w2g, b2g)
    sum((g1[l].float() * u1[l].to(cuda)).sum() for l in range(n)).backward()
    for l in range(n):
        err = float((w1n[l].grad.cpu() - w1r[l].grad).abs().max())
        assert err <= tol * float(w1r[l].grad.abs().max()) + 1e-5, (case, l, err)
