"""Parity of the AR-FPN kernels (gather, gated residual) and the AR-RFF gate
with the oracle (torch CPU ops = the reference arithmetic), fwd and bwd."""
import os

import numpy as np
import pytest
import torch

from util import assert_close_bf16, assert_close_fp32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FRCNN = [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5)]        # exact 2x chain
RAGGED = [(50, 84), (25, 42), (13, 21), (7, 11), (4, 6)]       # ceil-halving, non-integer ratios


def _inputs(oracle, shapes, batch=2, channels=8, seed=0):
    xs = oracle.synthetic_pyramid(batch, channels, shapes, seed=seed)
    gen = torch.Generator().manual_seed(seed + 100)
    hr, wr = shapes[2]
    bsf = torch.randn(batch, channels, hr, wr, generator=gen)
    g1 = [torch.randn(batch, 1, h, w, generator=gen) for h, w in shapes]
    g2 = [torch.randn(batch, 1, h, w, generator=gen) for h, w in shapes]
    return xs, bsf, g1, g2


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("shapes", [FRCNN, RAGGED])
@pytest.mark.parametrize("nhwc", [False, True])
def test_gather_forward_backward(oracle, cuda, shapes, nhwc):
    import arfe_b200 as A
    xs, _, _, _ = _inputs(oracle, shapes)
    xo = [x.clone().requires_grad_(True) for x in xs]
    ref = oracle.wfpn_gather(xo, 2)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    ref.backward(g)
    xg = [(_cl(x.to(cuda)) if nhwc else x.to(cuda)).requires_grad_(True) for x in xs]
    got = A.fpn_gather(xg, 2)
    # same op order as the reference (adds in level order, true division): exact
    assert torch.equal(got.cpu(), ref.detach()), float((got.cpu() - ref).abs().max())
    got.backward(g.to(cuda))
    for l in range(5):
        assert_close_fp32(xg[l].grad, xo[l].grad, f"gather grad level {l}")


@pytest.mark.parametrize("shapes", [FRCNN, RAGGED])
@pytest.mark.parametrize("nhwc", [False, True])
def test_apply_forward_backward(oracle, cuda, shapes, nhwc):
    import arfe_b200 as A
    xs, bsf, g1, g2 = _inputs(oracle, shapes, seed=3)
    req = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    xo, g1o, g2o, bo = req(xs), req(g1), req(g2), bsf.clone().requires_grad_(True)
    ref = oracle.wfpn_apply(xo, bo, g1o, g2o)
    gs = [torch.randn(r.shape, generator=torch.Generator().manual_seed(7 + i))
          for i, r in enumerate(ref)]
    torch.autograd.backward(ref, gs)
    mv = (lambda t: _cl(t.to(cuda))) if nhwc else (lambda t: t.to(cuda))
    xg = [mv(x).requires_grad_(True) for x in xs]
    bg = mv(bsf).requires_grad_(True)
    g1g = [t.to(cuda).requires_grad_(True) for t in g1]
    g2g = [t.to(cuda).requires_grad_(True) for t in g2]
    got = A.fpn_apply(xg, bg, g1g, g2g)
    for l in range(5):
        assert_close_fp32(got[l], ref[l], f"apply out level {l}")
    torch.autograd.backward(got, [g.to(cuda) for g in gs])
    for l in range(5):
        assert_close_fp32(xg[l].grad, xo[l].grad, f"apply dx level {l}")
        for name, a, b in (("dg1", g1g, g1o), ("dg2", g2g, g2o)):
            r = b[l].grad
            d = (a[l].grad.cpu() - r).abs().max()
            assert float(d) <= 2e-5 * float(r.abs().max()) + 1e-6, (name, l, float(d))
    d = (bg.grad.cpu() - bo.grad).abs().max()
    assert float(d) <= 2e-5 * float(bo.grad.abs().max()) + 1e-6


def test_arfpn_golden(cuda):
    import arfe_b200 as A
    d = np.load(os.path.join(GOLD, "arfpn_small.npz"))
    t = lambda k: torch.from_numpy(d[k]).to(cuda)
    xs = [t(f"x{l}") for l in range(5)]
    assert torch.equal(A.fpn_gather(xs, 2).cpu(), torch.from_numpy(d["gathered"]))
    outs = A.fpn_apply(xs, t("bsf"), [t(f"g1_{l}") for l in range(5)],
                       [t(f"g2_{l}") for l in range(5)])
    for l in range(5):
        assert_close_fp32(outs[l], torch.from_numpy(d[f"out{l}"]), f"golden out{l}")


def test_arfpn_bf16(oracle, cuda):
    import arfe_b200 as A
    xs, bsf, g1, g2 = _inputs(oracle, FRCNN, seed=9)
    b16 = lambda ts: [t.bfloat16() for t in ts]
    xb, g1b, g2b, bb = b16(xs), b16(g1), b16(g2), bsf.bfloat16()
    f32 = lambda ts: [t.float() for t in ts]
    ref_g = oracle.wfpn_gather(f32(xb), 2)
    ref_o = oracle.wfpn_apply(f32(xb), bb.float(), f32(g1b), f32(g2b))
    got_g = A.fpn_gather([x.to(cuda) for x in xb], 2)
    assert got_g.dtype == torch.bfloat16
    assert_close_bf16(got_g, ref_g, "gather bf16")
    got_o = A.fpn_apply([x.to(cuda) for x in xb], bb.to(cuda), [t.to(cuda) for t in g1b],
                        [t.to(cuda) for t in g2b])
    for l in range(5):
        assert_close_bf16(got_o[l], ref_o[l], f"apply bf16 level {l}")


def test_neck_module_matches_oracle_module(oracle, cuda):
    """Whole WFPNDualSpatial (convs + NonLocal2D on PyTorch, our kernels for
    gather/apply) against the oracle module with the same weights."""
    import arfe_b200 as A
    torch.manual_seed(0)
    ref_m = oracle.WFPNDualSpatial(16, 5)
    ref_m.init_weights()
    # make the zero-initialised conv_out non-trivial so the refine path matters
    torch.nn.init.normal_(ref_m.refine.conv_out.conv.weight, 0, 0.05)
    m = A.WFPNDualSpatial(16, 5)
    m.load_state_dict(ref_m.state_dict())
    xs = oracle.synthetic_pyramid(1, 16, FRCNN, seed=4)
    ref = ref_m(xs)
    got = m.to(cuda)([x.to(cuda) for x in xs])
    for l in range(5):
        # convs/matmuls run on cuDNN/cuBLAS (TF32 off by default for matmul; conv may differ)
        err = (got[l].cpu() - ref[l]).abs().max()
        assert float(err) <= 1e-3 * float(ref[l].abs().max()), (l, float(err))


@pytest.mark.parametrize("K,C", [(0, 8), (5, 8), (33, 16)])
def test_rff_gate(oracle, cuda, K, C):
    import arfe_b200 as A
    gen = torch.Generator().manual_seed(K + C)
    x = torch.randn(K, 3 * C, 7, 7, generator=gen)
    a = torch.randn(K, C, 7, 7, generator=gen).relu()
    b = torch.randn(K, C, 7, 7, generator=gen).relu()
    xo, ao, bo = (t.clone().requires_grad_(True) for t in (x, a, b))
    ref = oracle.rff_gate(xo[:, :C], ao, bo)
    g = torch.randn(ref.shape, generator=gen)
    xg, ag, bg = (t.to(cuda).requires_grad_(True) for t in (x, a, b))
    ori, _, _ = A.split3(xg, C)
    got = A.rff_gate(ori, ag, bg)
    assert_close_fp32(got, ref, "gate fwd")
    if K == 0:
        return
    ref.backward(g)
    got.backward(g.to(cuda))
    assert_close_fp32(xg.grad, xo.grad, "gate d x")
    assert_close_fp32(ag.grad, ao.grad, "gate d a")
    assert_close_fp32(bg.grad, bo.grad, "gate d b")
    # bf16
    gb = A.rff_gate(x[:, :C].bfloat16().to(cuda), a.bfloat16().to(cuda), b.bfloat16().to(cuda))
    refb = oracle.rff_gate(x[:, :C].bfloat16().float(), a.bfloat16().float(), b.bfloat16().float())
    assert_close_bf16(gb, refb, "gate bf16")


def test_head_and_roi_head_match_oracle(oracle, cuda):
    """MultiRoIsBBoxHead + StandardRoIHead._bbox_forward (AR-RFF enabled)."""
    import arfe_b200 as A
    from util import STRIDES, mixed_rois, small_pyramid
    torch.manual_seed(1)
    C = 16
    ref_h = oracle.MultiRoIsBBoxHead(in_channels=C, fc_out_channels=32, num_classes=3)
    ref_h.conv_out_channels = C
    rh = A.StandardRoIHead(
        bbox_roi_extractor=dict(type='SingleRoIExtractor',
                                roi_layer=dict(type='RoIAlign', out_size=7, sample_num=0),
                                out_channels=C, featmap_strides=list(STRIDES)),
        bbox_head=dict(type='MultiRoIsBBoxHead', in_channels=C, conv_out_channels=C,
                       fc_out_channels=32, roi_feat_size=7, num_classes=3))
    rh.bbox_head.load_state_dict(ref_h.state_dict())
    feats = small_pyramid(oracle, batch=2, channels=C)
    rois = mixed_rois(oracle, 40, 320, 192, 2, seed=2)
    bf = oracle.arrff_bbox_feats(feats, rois, list(STRIDES))
    cls_ref, reg_ref = ref_h(bf)
    res = rh.to(cuda)._bbox_forward([f.to(cuda) for f in feats], rois.to(cuda))
    assert_close_fp32(res["bbox_feats"], bf, "bbox_feats")
    for got, ref in ((res["cls_score"], cls_ref), (res["bbox_pred"], reg_ref)):
        err = (got.cpu() - ref).abs().max()
        assert float(err) <= 1e-3 * float(ref.abs().max()) + 1e-5


def test_roi_head_split_regions_forward_backward(oracle, cuda):
    """StandardRoIHead with the extractor's opt-in split mode (channels-last
    pyramid, regions as separate tensors straight into the head): scores and
    pyramid gradients against the oracle modules."""
    import arfe_b200 as A
    from util import STRIDES, mixed_rois, small_pyramid
    torch.manual_seed(2)
    C = 16
    ref_h = oracle.MultiRoIsBBoxHead(in_channels=C, fc_out_channels=32, num_classes=3)
    ref_h.conv_out_channels = C
    rh = A.StandardRoIHead(
        bbox_roi_extractor=dict(type='SingleRoIExtractor',
                                roi_layer=dict(type='RoIAlign', out_size=7, sample_num=0),
                                out_channels=C, featmap_strides=list(STRIDES)),
        bbox_head=dict(type='MultiRoIsBBoxHead', in_channels=C, conv_out_channels=C,
                       fc_out_channels=32, roi_feat_size=7, num_classes=3))
    rh.bbox_head.load_state_dict(ref_h.state_dict())
    rh.bbox_roi_extractor.roi_feats_split = True
    feats = small_pyramid(oracle, batch=2, channels=C)
    rois = mixed_rois(oracle, 40, 320, 192, 2, seed=5)
    fo = [f.clone().requires_grad_(True) for f in feats]
    cls_ref, reg_ref = ref_h(oracle.arrff_bbox_feats(fo, rois, list(STRIDES)))
    (cls_ref.sum() + reg_ref.square().sum()).backward()
    fg = [f.to(cuda).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
    res = rh.to(cuda)._bbox_forward(fg, rois.to(cuda))
    assert isinstance(res["bbox_feats"], tuple) and len(res["bbox_feats"]) == 3
    for got, ref in ((res["cls_score"], cls_ref), (res["bbox_pred"], reg_ref)):
        err = (got.detach().cpu() - ref.detach()).abs().max()
        assert float(err) <= 1e-3 * float(ref.abs().max()) + 1e-5
    (res["cls_score"].sum() + res["bbox_pred"].square().sum()).backward()
    for l in range(5):
        r = fo[l].grad if fo[l].grad is not None else torch.zeros_like(feats[l])
        err = (fg[l].grad.cpu() - r).abs().max()
        assert float(err) <= 1e-3 * float(r.abs().max()) + 1e-6, (l, float(err))
