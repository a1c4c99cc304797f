"""Randomised configurations of the fused extraction and the AR-FPN kernels
against the oracle: levels, channels (incl. counts the vector paths cannot
take), pool sizes, sampling ratio, regions, layout, batch -- forward and
backward.  Seeds are fixed; every case is a few milliseconds."""
import random

import pytest
import torch

from util import assert_close_fp32

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("seed", list(range(12)))
def test_roi_fuse_random_config(oracle, cuda, seed):
    import arfe_b200 as A
    rnd = random.Random(seed)
    L = rnd.choice([1, 2, 4, 5])
    strides = [4, 8, 16, 32, 64][:L]
    C = rnd.choice([3, 4, 8, 12, 20, 33])
    B = rnd.choice([1, 2, 3])
    out_size = rnd.choice([(7, 7), (3, 3), (14, 14), (2, 5), (1, 1)])
    sample_num = rnd.choice([0, 0, 2, 3])
    regions = rnd.choice([1, 3])
    layout_cl = rnd.choice([False, True])
    img_h, img_w = rnd.choice([(96, 160), (200, 136), (64, 64)])
    K = rnd.choice([1, 7, 40])
    shapes = oracle.pyramid_shapes(img_h, img_w, strides)
    feats = oracle.synthetic_pyramid(B, C, shapes, seed=seed)
    rois = oracle.synthetic_rois(K, img_w, img_h, B, seed=seed, smin=3.0, smax=float(min(img_h, img_w)))
    fo = [f.clone().requires_grad_(True) for f in feats]
    if regions == 3:
        ref = oracle.arrff_bbox_feats(fo, rois, strides, out_size=out_size, sample_num=sample_num)
    else:
        ref = oracle.single_roi_extractor(fo, rois, strides, out_size=out_size, sample_num=sample_num)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(seed + 1))
    ref.backward(g)
    mv = (lambda t: _cl(t.to(cuda))) if layout_cl else (lambda t: t.to(cuda))
    fg = [mv(f).requires_grad_(True) for f in feats]
    got = A.roi_fuse(fg, rois.to(cuda), out_size, [1.0 / s for s in strides], sample_num=sample_num,
                     regions=regions, out_channels_last=layout_cl)
    tag = f"seed={seed} L={L} C={C} B={B} out={out_size} sn={sample_num} R={regions} cl={layout_cl} K={K}"
    assert_close_fp32(got, ref, tag)
    got.backward(mv(g) if g.dim() == 4 else g.to(cuda))
    for l in range(L):
        r = fo[l].grad if fo[l].grad is not None else torch.zeros_like(feats[l])
        d = (fg[l].grad.cpu() - r).abs().max()
        assert float(d) <= 3e-5 * float(r.abs().max()) + 1e-6, (tag, l, float(d))


@pytest.mark.parametrize("seed", list(range(8)))
def test_fpn_random_config(oracle, cuda, seed):
    import arfe_b200 as A
    rnd = random.Random(100 + seed)
    C = rnd.choice([4, 6, 8, 16, 36])
    B = rnd.choice([1, 2])
    layout_cl = rnd.choice([False, True])
    base_h, base_w = rnd.choice([(48, 80), (50, 84), (36, 52), (64, 64)])
    L = rnd.choice([3, 4, 5])
    refine = rnd.choice([0, 1, 2]) if L > 2 else 0
    shapes, h, w = [], base_h, base_w
    for _ in range(L):
        shapes.append((h, w))
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    xs = oracle.synthetic_pyramid(B, C, shapes, seed=seed)
    gen = torch.Generator().manual_seed(seed + 7)
    hr, wr = shapes[refine]
    bsf = torch.randn(B, C, hr, wr, generator=gen)
    g1 = [torch.randn(B, 1, a, b, generator=gen) for a, b in shapes]
    g2 = [torch.randn(B, 1, a, b, generator=gen) for a, b in shapes]
    req = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    xo, g1o, g2o, bo = req(xs), req(g1), req(g2), bsf.clone().requires_grad_(True)
    gath_ref = oracle.wfpn_gather(xo, refine)
    outs_ref = oracle.wfpn_apply(xo, bo, g1o, g2o)
    gg = torch.randn(gath_ref.shape, generator=gen)
    gs = [torch.randn(o.shape, generator=gen) for o in outs_ref]
    torch.autograd.backward([gath_ref] + list(outs_ref), [gg] + gs)
    mv = (lambda t: _cl(t.to(cuda))) if layout_cl else (lambda t: t.to(cuda))
    xg = [mv(x).requires_grad_(True) for x in xs]
    bg = mv(bsf).requires_grad_(True)
    g1g = [t.to(cuda).requires_grad_(True) for t in g1]
    g2g = [t.to(cuda).requires_grad_(True) for t in g2]
    gath = A.fpn_gather(xg, refine)
    outs = A.fpn_apply(xg, bg, g1g, g2g)
    tag = f"seed={seed} C={C} B={B} cl={layout_cl} L={L} refine={refine} base={base_h}x{base_w}"
    assert torch.equal(gath.cpu(), gath_ref.detach()), tag
    for l in range(L):
        assert_close_fp32(outs[l], outs_ref[l], tag + f" out{l}")
    torch.autograd.backward([gath] + list(outs), [mv(gg)] + [mv(t) for t in gs])
    for l in range(L):
        assert_close_fp32(xg[l].grad, xo[l].grad, tag + f" dx{l}")
        for a, b in ((g1g, g1o), (g2g, g2o)):
            d = (a[l].grad.cpu() - b[l].grad).abs().max()
            assert float(d) <= 3e-5 * float(b[l].grad.abs().max()) + 1e-6, (tag, l)
    d = (bg.grad.cpu() - bo.grad).abs().max()
    assert float(d) <= 3e-5 * float(bo.grad.abs().max()) + 1e-6, tag
