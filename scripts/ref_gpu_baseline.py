#!/usr/bin/env python
"""GPU baseline: the REFERENCE's own CUDA RoIAlign v2 kernel (compiled unmodified
for sm_100a into oracle/_ref/cuda by `python oracle/build_oracle.py --cuda`)
driven by the reference's extractor logic -- per region a zero-filled output,
per level `inds = lvl == i; roi_feats[inds] = RoIAlign(feats[i], rois[inds])`
(single_level.py:116-149), three regions + torch.cat (standard_roi_head.py:138-155),
backward through autograd with the reference's backward_v2 -- on the bench
workload (2 x 800x1344, C=256, 1024 RoIs, NCHW fp32), next to our fused path.
MEASUREMENT INFRASTRUCTURE: nothing here is on the product path."""
import glob, importlib.util, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arfe_b200 import workload as wl, _lib as L          # noqa: E402
from oracle import arfe_oracle as O                      # noqa: E402

hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cuda", "roi_align_ref_cuda*.so"))
if not hits:
    print(json.dumps({"ref_gpu_baseline": "unavailable: oracle/_ref/cuda not built"})); sys.exit(0)
spec = importlib.util.spec_from_file_location("roi_align_ref_cuda", hits[0])
ext = importlib.util.module_from_spec(spec); spec.loader.exec_module(ext)


class RefRoIAlign(torch.autograd.Function):
    """RoIAlignFunction of the reference (ops/roi_align/roi_align.py:9-74), aligned=True."""
    @staticmethod
    def forward(ctx, feat, rois, scale):
        ctx.save_for_backward(rois); ctx.shape, ctx.scale = feat.shape, scale
        return ext.forward_v2(feat, rois, scale, 7, 7, 0, True)

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        B, C, H, W = ctx.shape
        return ext.backward_v2(g.contiguous(), rois, ctx.scale, 7, 7, B, C, H, W, 0, True), None, None


def extract(feats, rois, strides):
    """SingleRoIExtractor.forward (single_level.py:109-152)."""
    out = feats[0].new_zeros(rois.size(0), feats[0].size(1), 7, 7)
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvl = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=len(feats) - 1).long()
    for i in range(len(feats)):
        inds = lvl == i
        if inds.any():
            out[inds] = RefRoIAlign.apply(feats[i], rois[inds, :], 1.0 / strides[i])
    return out


def ref_step(feats, rois, strides, g):
    lh, lw = O.get_adaptive_scale_rois(rois, 1)
    F = torch.cat([extract(feats, rois, strides), extract(feats, lw, strides), extract(feats, lh, strides)], 1)
    if g is not None:
        F.backward(g)
    return F


def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


dev = torch.device("cuda:0")
host = wl.host_inputs(2, 512, 256)
strides = list(wl.STRIDES)
feats = [t.to(dev).requires_grad_(True) for t in host["x"]]
rois = host["rois"].to(dev)
g = torch.randn(rois.size(0), 768, 7, 7, device=dev)
with torch.no_grad():
    ref_fwd = timed(lambda: ref_step([f.detach() for f in feats], rois, strides, None))
def fb():
    for f in feats: f.grad = None
    ref_step(feats, rois, strides, g)
ref_fb = timed(fb)
# ours, same tensors' values, channels-last fast path through the C ABI
st = wl.TrainStep(wl.host_inputs(2, 512, 256, channels_last=True), dev)
ours_fwd = timed(lambda: L.check(st.roi_fuse_fwd(), "f"), n=30)
def ours():
    L.check(st.roi_fuse_fwd(), "f"); L.check(st.roi_fuse_bwd(), "b")
ours_fb = timed(ours, n=30)
# values agree?
with torch.no_grad():
    Fr = ref_step([f.detach() for f in feats], rois, strides, None)
import arfe_b200 as A
Fo = A.roi_fuse([f.detach() for f in feats], rois, 7, [1.0 / s for s in strides], regions=3)
err = float((Fr - Fo).abs().max() / Fr.abs().max())
print(json.dumps({"workload": "2 x 800x1344, C=256, 1024 RoIs x 3 regions, 7x7, fp32",
                  "reference_cuda_v2_extraction_ms": {"fwd": round(ref_fwd, 3), "fwd+bwd": round(ref_fb, 3)},
                  "ours_extraction_ms": {"fwd": round(ours_fwd, 3), "fwd+bwd": round(ours_fb, 3)},
                  "speedup": {"fwd": round(ref_fwd / ours_fwd, 1), "fwd+bwd": round(ref_fb / ours_fb, 1)},
                  "max_rel_diff_fwd": err}))
