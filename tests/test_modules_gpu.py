"""Module-surface cases the round-1 review found untested: the cascade's per-stage
AR-RFF block (cascade_roi_head.py:120-142), the extractor's fp16 contract
(@force_fp32(apply_to=('feats',), out_fp16=True), single_level.py:109), its hook
arguments in the reference's order of operations (single_level.py:126-152), the gate
on a single channels-last RoI, and reuse of a built RoI plan by several forwards."""
import pytest
import torch

from util import STRIDES, assert_close_fp32, mixed_rois, small_pyramid

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _extractor_cfg(C, strides):
    return dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', out_size=7, sample_num=0),
                out_channels=C, featmap_strides=list(strides))


@pytest.mark.parametrize("split", [False, True])
def test_cascade_roi_head_three_stages(oracle, cuda, split):
    """CascadeRoIHead._bbox_forward(stage, x, rois) for the three stages with boxes that
    change between stages (the refinement itself is the caller's, cascade_roi_head.py:299):
    scores per stage and the pyramid gradient of all stages against the oracle modules."""
    import arfe_b200 as A
    torch.manual_seed(3)
    C, strides = 16, STRIDES[:4]
    head_cfg = dict(type='MultiRoIsBBoxHead', in_channels=C, conv_out_channels=C,
                    fc_out_channels=32, roi_feat_size=7, num_classes=3)
    rh = A.CascadeRoIHead(3, _extractor_cfg(C, strides), head_cfg)
    ref_heads = []
    for s in range(3):
        h = oracle.MultiRoIsBBoxHead(in_channels=C, fc_out_channels=32, num_classes=3)
        h.conv_out_channels = C
        for p in h.parameters():
            torch.nn.init.normal_(p, 0, 0.05)
        rh.bbox_head[s].load_state_dict(h.state_dict())
        rh.bbox_roi_extractor[s].roi_feats_split = split
        ref_heads.append(h)
    feats = small_pyramid(oracle, batch=2, channels=C, strides=strides)
    rois = [mixed_rois(oracle, 30, 320, 192, 2, seed=10 + s) for s in range(3)]
    fo = [f.clone().requires_grad_(True) for f in feats]
    loss_ref, ref = 0, []
    for s in range(3):
        cls, reg = ref_heads[s](oracle.arrff_bbox_feats(fo, rois[s], list(strides)))
        ref.append((cls, reg))
        loss_ref = loss_ref + (0.5 + s) * (cls.sum() + reg.square().sum())
    loss_ref.backward()
    fg = [(_cl(f.to(cuda)) if split else f.to(cuda)).requires_grad_(True) for f in feats]
    rh.to(cuda)
    loss = 0
    for s in range(3):
        res = rh._bbox_forward(s, fg, rois[s].to(cuda))
        if split:
            assert isinstance(res["bbox_feats"], tuple) and len(res["bbox_feats"]) == 3
        for got, want in ((res["cls_score"], ref[s][0]), (res["bbox_pred"], ref[s][1])):
            err = (got.detach().cpu() - want.detach()).abs().max()
            assert float(err) <= 1e-3 * float(want.abs().max()) + 1e-5, (s, float(err))
        loss = loss + (0.5 + s) * (res["cls_score"].sum() + res["bbox_pred"].square().sum())
    loss.backward()
    for l in range(4):
        r = fo[l].grad if fo[l].grad is not None else torch.zeros_like(feats[l])
        err = (fg[l].grad.cpu() - r).abs().max()
        assert float(err) <= 1e-3 * float(r.abs().max()) + 1e-6, (l, float(err))


def test_extractor_fp16_contract(oracle, cuda):
    """fp16 pyramid in -> fp32 math -> fp16 RoI features out (force_fp32 ... out_fp16=True)."""
    import arfe_b200 as A
    C = 16
    ext = A.SingleRoIExtractor(dict(type='RoIAlign', out_size=7, sample_num=0), C, list(STRIDES[:4]))
    feats = [f.half() for f in small_pyramid(oracle, batch=2, channels=C, strides=STRIDES[:4])]
    rois = mixed_rois(oracle, 25, 320, 192, 2, seed=4)
    f32 = [f.float() for f in feats]
    want1 = oracle.single_roi_extractor(f32, rois, list(STRIDES[:4]))
    want3 = oracle.arrff_bbox_feats(f32, rois, list(STRIDES[:4]))
    fg = [f.to(cuda) for f in feats]
    got1 = ext(fg, rois.to(cuda))
    got3 = ext.forward_regions(fg, rois.to(cuda), regions=3, facs=1)
    assert got1.dtype == torch.float16 and got3.dtype == torch.float16
    for got, want in ((got1, want1), (got3, want3)):
        # the fp32 result is within 1e-5 of the oracle's; rounding both to fp16 can differ by one fp16 ulp
        err = (got.float().cpu() - want.half().float()).abs()
        assert bool((err <= 2.0 ** -10 * want.abs() + 1e-6).all()), float(err.max())
    ext.roi_feats_split = True
    parts = ext.forward_regions([_cl(f) for f in fg], rois.to(cuda), regions=3, facs=1)
    assert all(p.dtype == torch.float16 for p in parts)
    # another kernel (channels-last ring) than got3's: fp32 results agree to ~1e-7, so the
    # fp16 roundings may differ by one ulp
    err = (torch.cat(parts, 1).float() - got3.float()).abs()
    assert bool((err <= 2.0 ** -10 * got3.float().abs() + 1e-6).all()), float(err.max())


def test_extractor_hooks_follow_reference_order(oracle, cuda):
    """roi_scale_factor / lvl / replace_rois: levels come from the ORIGINAL rois (or
    replace_rois), shifted by lvl; the rescaled boxes are only what is sampled."""
    import arfe_b200 as A
    C, strides = 8, list(STRIDES[:4])
    ext = A.SingleRoIExtractor(dict(type='RoIAlign', out_size=7, sample_num=2), C, strides)
    feats = small_pyramid(oracle, batch=2, channels=C, strides=strides)
    # sides well inside the level bands (56 * 2^k thresholds), so the level of a box
    # moves when it is rescaled by 1.6 but not through rounding
    rows = []
    for i, side in enumerate([40.0, 80.0, 100.0, 150.0, 180.0, 75.0, 30.0, 160.0]):
        x0, y0 = 5.0 + 9 * i, 3.0 + 4 * i
        rows.append([i % 2, x0, y0, x0 + side, y0 + side * 0.9])
    rois = torch.tensor(rows, dtype=torch.float32)
    repl = rois.clone()
    repl[:, 3:] = repl[:, 1:3] + (repl[:, 3:] - repl[:, 1:3]) * 0.5

    def want(roi_scale_factor=None, lvl=None, replace_rois=None):
        tgt = oracle.map_roi_levels(replace_rois if replace_rois is not None else rois, 4)
        if lvl is not None:
            tgt = (tgt + lvl).clamp(0, 3)
        r = ext.roi_rescale(rois, roi_scale_factor) if roi_scale_factor is not None else rois
        out = torch.zeros(len(rois), C, 7, 7)
        for i in range(4):
            m = tgt == i
            if m.any():
                out[m] = oracle.roi_align_forward(feats[i], r[m], 7, 1 / strides[i], 2)
        return out, tgt

    fg = [f.to(cuda) for f in feats]
    base_lv = oracle.map_roi_levels(rois, 4)
    scaled_lv = oracle.map_roi_levels(ext.roi_rescale(rois, 1.6), 4)
    assert not torch.equal(base_lv, scaled_lv), "the case must distinguish the two orders"
    for kw in (dict(roi_scale_factor=1.6), dict(lvl=1), dict(lvl=-1), dict(replace_rois=repl),
               dict(roi_scale_factor=0.7, lvl=1, replace_rois=repl)):
        w, _ = want(**kw)
        gkw = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in kw.items()}
        assert_close_fp32(ext(fg, rois.to(cuda), **gkw), w, str(sorted(kw)))


@pytest.mark.parametrize("K", [1, 2])
@pytest.mark.parametrize("source", ["split", "cat_cl_slice", "cat_nchw_slice"])
def test_gate_single_roi_every_layout(oracle, cuda, K, source):
    """K == 1 with a channels-last `ori` (a region tensor of roi_fuse_split, or a
    channel slice of the channels-last cat tensor) -- forward and both gradients."""
    import arfe_b200 as A
    C = 16
    gen = torch.Generator().manual_seed(K)
    x = torch.randn(K, 3 * C, 7, 7, generator=gen)
    a = torch.randn(K, C, 7, 7, generator=gen).relu()
    b = torch.randn(K, C, 7, 7, generator=gen).relu()
    g = torch.randn(K, C, 7, 7, generator=gen)
    xo, ao, bo = (t.clone().requires_grad_(True) for t in (x, a, b))
    ref = oracle.rff_gate(xo[:, :C], ao, bo)
    ref.backward(g)
    if source == "split":
        xg = _cl(x[:, :C].contiguous().to(cuda)).requires_grad_(True)
        ori = xg
    else:
        xg = (_cl(x.to(cuda)) if source == "cat_cl_slice" else x.to(cuda)).requires_grad_(True)
        ori = xg[:, :C]
    # a, b arrive in whatever layout the convs produced: try the "other" one
    ag = (a.to(cuda) if source != "cat_nchw_slice" else _cl(a.to(cuda))).requires_grad_(True)
    bg = _cl(b.to(cuda)).requires_grad_(True)
    got = A.rff_gate(ori, ag, bg)
    assert_close_fp32(got, ref, f"gate fwd K={K} {source}")
    got.backward(g.to(cuda))
    want_dx = xo.grad[:, :C] if source == "split" else xo.grad
    assert_close_fp32(xg.grad, want_dx, "gate d ori")
    assert_close_fp32(ag.grad, ao.grad, "gate d a")
    assert_close_fp32(bg.grad, bo.grad, "gate d b")
    # mixed precision: bf16 a / b with an fp32 ori are converted, not reinterpreted
    got2 = A.rff_gate(ori.detach(), a.bfloat16().to(cuda), b.bfloat16().to(cuda))
    ref2 = oracle.rff_gate(x[:, :C], a.bfloat16().float(), b.bfloat16().float())
    assert_close_fp32(got2, ref2, "gate with bf16 a/b")


def test_forward_reuses_a_built_plan(oracle, cuda):
    """arfe_roi_fuse_forward_plan(_split) with plan_ready = 1 must serve any
    number of calls on the same plan (fixed RoIs over iterations, a captured forward)."""
    from arfe_b200 import _lib as L
    lib = L.lib()
    C, strides = 64, list(STRIDES[:4])
    feats = [_cl(f.to(cuda)) for f in small_pyramid(oracle, batch=2, channels=C, strides=strides)]
    rois = mixed_rois(oracle, 40, 320, 192, 2, seed=8).to(cuda)
    K = rois.shape[0]
    Hs, Ws = L.int_array([f.shape[2] for f in feats]), L.int_array([f.shape[3] for f in feats])
    scales = L.float_array([1.0 / s for s in strides])
    nbytes = lib.arfe_roi_plan_bytes(K, 3, 4, 2, Hs, Ws)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=cuda)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    stream = L.stream_ptr(cuda)
    geo = (Hs, Ws, scales, 4, 2, C, rois.data_ptr(), K, 3, 1.0, 7, 7, 0, 56.0, L.ARFE_F32)
    L.check(lib.arfe_roi_plan_build(*geo, ws_ptr, nbytes, stream), "plan")
    want = oracle.arrff_bbox_feats([f.cpu().contiguous() for f in feats], rois.cpu(), strides)
    outs = []
    for it in range(3):
        out = torch.full((K, 3 * C, 7, 7), float("nan"), device=cuda).contiguous(memory_format=torch.channels_last)
        L.check(lib.arfe_roi_fuse_forward_plan(
            L.ptr_array(feats), Hs, Ws, scales, 4, 2, C, rois.data_ptr(), K, 3, 1.0, 7, 7, 0, 56.0,
            L.ARFE_F32, out.data_ptr(), None, None, ws_ptr, nbytes, 1, stream), "forward on a ready plan")
        torch.cuda.synchronize()
        assert_close_fp32(out, want, f"forward #{it} on the same plan")
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    parts = [torch.full((K, C, 7, 7), float("nan"), device=cuda).contiguous(memory_format=torch.channels_last)
             for _ in range(3)]
    for it in range(2):
        L.check(lib.arfe_roi_fuse_forward_plan_split(
            L.ptr_array(feats), Hs, Ws, scales, 4, 2, C, rois.data_ptr(), K, 3, 1.0, 7, 7, 0, 56.0,
            L.ARFE_F32, L.ptr_array(parts), ws_ptr, nbytes, 1, stream), "split forward on a ready plan")
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(parts, 1), outs[0])


def test_use_torchvision_is_refused():
    import arfe_b200 as A
    with pytest.raises(NotImplementedError):
        A.RoIAlign(7, 0.25, use_torchvision=True, aligned=False)


@pytest.mark.parametrize("layout", ["nchw_cat", "cl_cat", "cl_split"])
@pytest.mark.parametrize("K,C", [(0, 8), (1, 16), (37, 20), (64, 256)])
def test_rff_softmax_fusion_variant(oracle, cuda, layout, K, C):
    """the softmax-over-regions fusion (multirois_bbox_head.py:187-197,
    commented in the shipped tree) -- forward and all four gradients against the oracle's
    restatement, for the concatenated NCHW / channels-last tensor and for split region tensors."""
    import arfe_b200 as A
    gen = torch.Generator().manual_seed(K * 7 + C)
    x = torch.randn(K, 3 * C, 7, 7, generator=gen)
    logits = torch.randn(K, 3, 7, 7, generator=gen) * 2
    g = torch.randn(K, C, 7, 7, generator=gen)
    xo, lo = x.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    ref = oracle.rff_softmax_fuse((xo[:, :C], xo[:, C:2 * C], xo[:, 2 * C:]), lo)
    if layout == "cl_split":
        parts = [_cl(x[:, j * C:(j + 1) * C].contiguous().to(cuda)).requires_grad_(True) for j in range(3)]
        regions = tuple(parts)
    else:
        xg = (_cl(x.to(cuda)) if layout == "cl_cat" else x.to(cuda)).requires_grad_(True)
        regions = xg
    lg = logits.to(cuda).requires_grad_(True)
    got = A.rff_softmax_fuse(regions, lg)
    assert_close_fp32(got, ref, f"softmax fusion fwd {layout}")
    if K == 0:
        return
    ref.backward(g)
    got.backward(g.to(cuda))
    dx = torch.cat([p.grad for p in parts], 1) if layout == "cl_split" else xg.grad
    assert_close_fp32(dx, xo.grad, "softmax fusion d regions")
    err = (lg.grad.cpu() - lo.grad).abs().max()
    assert float(err) <= 1e-5 * float(lo.grad.abs().max()) + 1e-5, float(err)   # a 3*C-term sum per bin
    # bf16 I/O
    gb = A.rff_softmax_fuse(x.bfloat16().to(cuda), logits.bfloat16().to(cuda))
    xb = x.bfloat16().float()
    refb = oracle.rff_softmax_fuse((xb[:, :C], xb[:, C:2 * C], xb[:, 2 * C:]), logits.bfloat16().float())
    err = (gb.float().cpu() - refb).abs()
    assert bool((err <= 1e-2 * refb.abs() + 1e-2 * float(refb.abs().mean())).all())


@pytest.mark.parametrize("n", [1, 63, 64, 65, 700, 5000])
def test_nms_matches_reference(oracle, cuda, n):
    """arfe_nms (device-side sweep) keeps exactly the boxes the reference's
    NMS keeps (its own nms_ext when oracle/_ref/nms travelled, else the pinned restatement),
    clustered boxes, several thresholds."""
    import arfe_b200 as A
    gen = torch.Generator().manual_seed(n)
    ncl = max(n // 12, 1)
    centres = torch.rand(ncl, 2, generator=gen) * torch.tensor([1300.0, 780.0])
    ctr = centres[torch.randint(0, ncl, (n,), generator=gen)] + torch.randn(n, 2, generator=gen) * 12
    wh = torch.exp(torch.randn(n, 2, generator=gen) * 0.4) * 60
    dets = torch.cat([ctr - wh / 2, ctr + wh / 2, torch.rand(n, 1, generator=gen)], 1).contiguous()
    backend = "ref" if oracle.ref_nms_ext() is not None else "py"
    for thr in (0.3, 0.5, 0.7):
        want = oracle.nms(dets, thr, backend)
        got_dets, got = A.nms(dets.to(cuda), thr)
        assert torch.equal(got.cpu(), want), (n, thr, len(want), len(got))
        assert torch.equal(got_dets.cpu(), dets[want])


def test_batched_nms_and_bbox2roi(oracle, cuda):
    import arfe_b200 as A
    gen = torch.Generator().manual_seed(5)
    n = 900
    ctr = torch.rand(n, 2, generator=gen) * 400
    wh = torch.rand(n, 2, generator=gen) * 90 + 4
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    scores = torch.rand(n, generator=gen)
    lvl = torch.randint(0, 5, (n,), generator=gen)
    dets, keep = A.batched_nms(boxes.to(cuda), scores.to(cuda), lvl.to(cuda), dict(type='nms', iou_thr=0.7))
    # reference composition (nms_wrapper.py:143-157) on the oracle
    off = lvl.float() * (boxes.max() + 1)
    want = oracle.nms(torch.cat([boxes + off[:, None], scores[:, None]], 1), 0.7, "py")
    assert torch.equal(keep.cpu(), want)
    assert torch.equal(dets.cpu(), torch.cat([boxes[want], scores[want, None]], 1))
    # proposals -> RoIs, image-major
    lists = [dets[:300, :5], dets[:0], dets[300:450, :4].contiguous(), dets[450:]]
    rois = A.bbox2roi(lists)
    assert torch.equal(rois.cpu(), oracle.bbox2roi([t.cpu() for t in lists]))
    assert A.bbox2roi([dets[:0], dets[:0]]).shape == (0, 5)


@pytest.mark.parametrize("nhwc", [False, True])
def test_neck_module_backward_with_linked_gradients(oracle, cuda, nhwc):
    """WFPNDualSpatial forward + backward through the module (gather and gated residual coupled
    by an FPNLink: d x_l = d out_l + gather gradient written once) against the oracle module with
    the same weights: input gradients and every parameter gradient."""
    import arfe_b200 as A
    torch.manual_seed(0)
    shapes = [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5)]
    ref_m = oracle.WFPNDualSpatial(16, 5)
    ref_m.init_weights()
    torch.nn.init.normal_(ref_m.refine.conv_out.conv.weight, 0, 0.05)
    m = A.WFPNDualSpatial(16, 5)
    m.load_state_dict(ref_m.state_dict())
    m = m.to(cuda)
    if nhwc:
        m = m.to(memory_format=torch.channels_last)
    xs = oracle.synthetic_pyramid(2, 16, shapes, seed=4)
    gs = [torch.randn(x.shape, generator=torch.Generator().manual_seed(30 + i)) for i, x in enumerate(xs)]
    xo = [x.clone().requires_grad_(True) for x in xs]
    torch.autograd.backward(list(ref_m(xo)), gs)
    xg = [(_cl(x.to(cuda)) if nhwc else x.to(cuda)).requires_grad_(True) for x in xs]
    out = m(xg)
    torch.autograd.backward(list(out), [g.to(cuda) for g in gs])
    for l in range(5):
        err = (xg[l].grad.cpu() - xo[l].grad).abs().max()
        assert float(err) <= 1e-3 * float(xo[l].grad.abs().max()) + 1e-6, (l, float(err))
    ref_p = dict(ref_m.named_parameters())
    for name, prm in m.named_parameters():
        r = ref_p[name].grad
        assert prm.grad is not None, name
        err = (prm.grad.cpu() - r).abs().max()
        # (refine.phi's bias shifts every logit of a softmax row alike: its gradient is exactly 0
        # in exact arithmetic, ~1e-6 of rounding noise on both sides -- hence the absolute floor)
        assert float(err) <= 2e-3 * float(r.abs().max()) + 1e-5, (name, float(err))
