"""CPU suite: pins the oracle.

* plain-C restatement == the reference's own RoIAlign compiled unmodified
  (oracle/_ref) == torchvision, bit for bit, forward and backward;
* oracle == the committed golden vectors (tests/golden, minted from oracle/_ref);
* the device-side level rule (csrc/geometry.cuh floor_log2_rn), restated here
  in numpy, == torch CPU floor(log2(.)) around every threshold and on a
  random sweep.
"""
import os

import numpy as np
import pytest
import torch

from util import STRIDES, mixed_rois, small_pyramid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_or_skip(oracle):
    if oracle.ref_ext() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference once)")


def test_c_oracle_equals_compiled_reference_and_torchvision(oracle):
    import torchvision
    _ref_or_skip(oracle)
    feats = small_pyramid(oracle, batch=2, channels=6)
    rois = mixed_rois(oracle, 60, 320, 192, 2, seed=1)
    for lvl, s in enumerate(STRIDES):
        for out_size, sn in (((7, 7), 0), ((3, 5), 2), ((14, 14), 0)):
            a = oracle.roi_align_forward(feats[lvl], rois, out_size, 1 / s, sn, "c")
            b = oracle.roi_align_forward(feats[lvl], rois, out_size, 1 / s, sn, "ref")
            c = torchvision.ops.roi_align(feats[lvl], rois, out_size, 1 / s, sn, True)
            assert torch.equal(a, b) and torch.equal(b, c), (lvl, out_size, sn)
            g = torch.randn(a.shape, generator=torch.Generator().manual_seed(lvl))
            ga = oracle.roi_align_backward(g, rois, out_size, 1 / s, feats[lvl].shape, sn, "c")
            gb = oracle.roi_align_backward(g, rois, out_size, 1 / s, feats[lvl].shape, sn, "ref")
            assert torch.equal(ga, gb), (lvl, out_size, sn)


def test_oracle_matches_golden_arrff(oracle):
    d = np.load(os.path.join(GOLD, "arrff_small.npz"))
    feats = [torch.from_numpy(d[f"feat{l}"]) for l in range(5)]
    rois = torch.from_numpy(d["rois"])
    boxes, lvls = oracle.region_boxes_and_levels(rois, 5)
    assert np.array_equal(boxes.numpy(), d["boxes"])
    assert np.array_equal(lvls.numpy().astype(np.int32), d["lvls"])
    fo = [f.clone().requires_grad_(True) for f in feats]
    out = oracle.arrff_bbox_feats(fo, rois, list(STRIDES), backend="c")
    assert np.array_equal(out.detach().numpy(), d["out"])
    out.backward(torch.from_numpy(d["grad_out"]))
    for l in range(5):
        g = fo[l].grad if fo[l].grad is not None else torch.zeros_like(feats[l])
        assert np.array_equal(g.numpy(), d[f"dfeat{l}"]), l


def test_oracle_matches_golden_roi_align_op(oracle):
    d = np.load(os.path.join(GOLD, "roi_align_op.npz"))
    feat, rois = torch.from_numpy(d["feat"]), torch.from_numpy(d["rois"])
    for sn in (0, 2):
        o = oracle.roi_align_forward(feat, rois, 3, 1 / 8, sn, "c")
        assert np.array_equal(o.numpy(), d[f"out_sn{sn}"])
        gi = oracle.roi_align_backward(torch.from_numpy(d[f"gout_sn{sn}"]), rois, 3, 1 / 8,
                                       feat.shape, sn, "c")
        assert np.array_equal(gi.numpy(), d[f"gin_sn{sn}"])


def test_oracle_matches_golden_arfpn(oracle):
    d = np.load(os.path.join(GOLD, "arfpn_small.npz"))
    xs = [torch.from_numpy(d[f"x{l}"]) for l in range(5)]
    g1 = [torch.from_numpy(d[f"g1_{l}"]) for l in range(5)]
    g2 = [torch.from_numpy(d[f"g2_{l}"]) for l in range(5)]
    assert np.array_equal(oracle.wfpn_gather(xs, 2).numpy(), d["gathered"])
    outs = oracle.wfpn_apply(xs, torch.from_numpy(d["bsf"]), g1, g2)
    for l in range(5):
        np.testing.assert_allclose(outs[l].numpy(), d[f"out{l}"], rtol=1e-6, atol=1e-7)


def test_product_region_generator_equals_oracle(oracle):
    """arfe_b200.get_adaptive_scale_rois (host/torch form) is bit-identical to
    the reference expressions (additional.py:38-71)."""
    from arfe_b200 import get_adaptive_scale_rois
    rois = mixed_rois(oracle, 500, 1344, 800, 2, seed=9)
    ah, aw = get_adaptive_scale_rois(rois, 1)
    bh, bw = oracle.get_adaptive_scale_rois(rois, 1)
    assert torch.equal(ah, bh) and torch.equal(aw, bw)


# ---- the device level rule, restated (csrc/geometry.cuh: floor_log2_rn) ----
def _floor_log2_rn(v):
    bits = v.view(np.uint32)
    e = ((bits >> 23) & 0xFF).astype(np.int64) - 127
    m = (bits & 0x7FFFFF).astype(np.int64)
    k = e + 1
    j = np.where(k >= 5, 2, np.where(k >= 3, 1, 0))
    return np.where(m + j >= 0x800000, k, e)


def _device_level(v, L):
    out = np.empty(v.shape, dtype=np.int64)
    nan = np.isnan(v)
    big = v >= 256.0
    small = v < 1.0
    mid = ~(nan | big | small)
    out[nan] = -1
    out[big] = L - 1
    out[small] = 0
    out[mid] = np.clip(_floor_log2_rn(v[mid]), 0, L - 1)
    return out


@pytest.mark.parametrize("L", [2, 4, 5, 8])
def test_level_rule_matches_torch_cpu(L):
    vs = []
    for k in range(0, 9):
        c = np.float32(2.0 ** k).view(np.int32)
        vs.append(np.arange(c - 4096, c + 4096, dtype=np.int32).view(np.float32))
    rng = np.random.default_rng(0)
    vs.append(np.exp(rng.uniform(np.log(1e-6), np.log(600.0), 2_000_000)).astype(np.float32))
    v = np.concatenate(vs)
    ref = torch.floor(torch.log2(torch.from_numpy(v.copy()))).clamp(min=0, max=L - 1).long().numpy()
    got = _device_level(v, L)
    assert np.array_equal(got, ref), np.flatnonzero(got != ref)[:10]


def test_level_thresholds_golden(oracle):
    d = np.load(os.path.join(GOLD, "level_thresholds.npz"))
    rois = torch.from_numpy(d["rois"])
    assert np.array_equal(oracle.map_roi_levels(rois, 5).numpy().astype(np.int32), d["lvls"])
    # and the restated device rule on the same boxes (exact fp32 steps)
    w = (rois[:, 3] - rois[:, 1]).numpy()
    h = (rois[:, 4] - rois[:, 2]).numpy()
    scale = np.sqrt((w * h).astype(np.float32)).astype(np.float32)
    v = (scale / np.float32(56.0)).astype(np.float32) + np.float32(1e-6)
    assert np.array_equal(_device_level(v.astype(np.float32), 5), d["lvls"])


def test_oracle_nms_matches_reference_and_golden(oracle):
    """The numpy restatement of nms_cpu.cpp == the reference's own nms_ext compiled unmodified
    (when oracle/_ref/nms travelled) == the committed fixture minted from it."""
    d = np.load(os.path.join(GOLD, "nms_small.npz"))
    dets = torch.from_numpy(d["dets"])
    for thr in (0.3, 0.5, 0.7):
        keep = oracle.nms(dets, thr, "py")
        assert np.array_equal(keep.numpy(), d[f"keep_{int(thr * 10)}"]), thr
    if oracle.ref_nms_ext() is not None:
        gen = torch.Generator().manual_seed(3)
        for n in (1, 2, 65, 257):
            ctr = torch.rand(n, 2, generator=gen) * 100
            wh = torch.rand(n, 2, generator=gen) * 50 + 1
            x = torch.cat([ctr - wh / 2, ctr + wh / 2, torch.rand(n, 1, generator=gen)], 1)
            for thr in (0.1, 0.5, 0.9):
                assert torch.equal(oracle.nms(x, thr, "py"), oracle.nms(x, thr, "ref")), (n, thr)
    lists = [torch.rand(3, 5), torch.zeros(0, 5), torch.rand(2, 4)]
    r = oracle.bbox2roi(lists)
    assert r.shape == (5, 5) and r[:, 0].tolist() == [0, 0, 0, 2, 2]


def test_nonlocal_attention_golden_and_module(oracle):
    """The restated NonLocal2D attention (non_local.py:65-101) reproduces its committed fixture,
    equals a literal transcription of the reference's forward, and rounding the operands to bf16
    moves it by no more than the bound the tensor-core kernel is tested against."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nonlocal_small.npz"))
    th, ph, gx = (torch.from_numpy(z[k]) for k in ("theta", "phi", "g"))
    assert torch.equal(oracle.nonlocal_attention(th, ph, gx), torch.from_numpy(z["y"]))
    assert torch.equal(oracle.nonlocal_attention(th, ph, gx, use_scale=True), torch.from_numpy(z["y_scaled"]))
    n, c, h, w = th.shape
    g_x = gx.view(n, c, -1).permute(0, 2, 1)
    theta_x = th.view(n, c, -1).permute(0, 2, 1)
    phi_x = ph.view(n, c, -1)
    pw = torch.matmul(theta_x, phi_x)
    pw /= theta_x.shape[-1] ** 0.5
    pw = pw.softmax(dim=-1)
    y = torch.matmul(pw, g_x).permute(0, 2, 1).contiguous().reshape(n, c, h, w)
    assert torch.equal(y, torch.from_numpy(z["y_scaled"]))
    yb = oracle.nonlocal_attention(th, ph, gx, round_operands=torch.bfloat16)
    assert (yb - torch.from_numpy(z["y"])).abs().max() <= 1e-2 * torch.from_numpy(z["y"]).abs().max()
