"""Full-size parity against the REFERENCE'S OWN CUDA kernel: roi_align_kernel_v2.cu
compiled unmodified for sm_100a (oracle/_ref/cuda, built by
`python oracle/build_oracle.py --cuda`; test infrastructure) and driven by the
reference's extractor logic (single_level.py:109-152, standard_roi_head.py:138-155)
on the BASELINE configs[1] workload -- 2 x 800x1344, C = 256, 1024 RoIs, 3 regions --
where the CPU oracle is too slow.  Two fp32 implementations with different
summation orders are compared, so the bound is 2e-5 of the tensor's scale (the
north-star bound of 1e-5 relative is asserted against the oracle in the other
tests); the reference's backward uses atomicAdd (run-to-run rounding noise)."""
import glob
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_ext():
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "cuda", "roi_align_ref_cuda*.so"))
    if not hits:
        pytest.skip("oracle/_ref/cuda not built (python oracle/build_oracle.py --cuda)")
    spec = importlib.util.spec_from_file_location("roi_align_ref_cuda", hits[0])
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    return ext


def test_full_size_forward_backward_vs_reference_cuda_kernel(oracle, cuda):
    import arfe_b200 as A
    from arfe_b200 import workload as wl
    ext = _ref_ext()
    strides = list(wl.STRIDES)

    class RefRoIAlign(torch.autograd.Function):  # ops/roi_align/roi_align.py:9-74, aligned=True
        @staticmethod
        def forward(ctx, feat, rois, scale):
            ctx.save_for_backward(rois)
            ctx.shape, ctx.scale = feat.shape, scale
            return ext.forward_v2(feat, rois, scale, 7, 7, 0, True)

        @staticmethod
        def backward(ctx, g):
            (rois,) = ctx.saved_tensors
            B, C, H, W = ctx.shape
            return ext.backward_v2(g.contiguous(), rois, ctx.scale, 7, 7, B, C, H, W, 0, True), None, None

    def extract(feats, rois):
        out = feats[0].new_zeros(rois.size(0), feats[0].size(1), 7, 7)
        lvl = oracle.map_roi_levels(rois.cpu(), len(feats)).to(rois.device)
        for i in range(len(feats)):
            inds = lvl == i
            if inds.any():
                out[inds] = RefRoIAlign.apply(feats[i], rois[inds, :], 1.0 / strides[i])
        return out

    host = wl.host_inputs(2, 512, 256, seed=11)
    rois = host["rois"].to(cuda)
    lh, lw = oracle.get_adaptive_scale_rois(host["rois"], 1)
    fr = [t.to(cuda).requires_grad_(True) for t in host["x"]]
    ref = torch.cat([extract(fr, rois), extract(fr, lw.to(cuda)), extract(fr, lh.to(cuda))], 1)
    g = torch.randn(ref.shape, device=cuda, generator=torch.Generator(device=cuda).manual_seed(5))
    ref.backward(g)

    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    fo = [cl(t.to(cuda)).requires_grad_(True) for t in host["x"]]
    parts = A.roi_fuse_split(fo, rois, 7, [1.0 / s for s in strides], regions=3)
    got = torch.cat(parts, 1)
    scale = float(ref.detach().abs().max())
    diff = float((got.detach() - ref.detach()).abs().max())
    assert diff <= 2e-5 * scale, diff / scale
    C = 256
    torch.autograd.backward(parts, [cl(g[:, r * C:(r + 1) * C]) for r in range(3)])
    for l in range(5):
        gs = float(fr[l].grad.abs().max())
        d = float((fo[l].grad - fr[l].grad).abs().max())
        assert d <= 5e-5 * gs + 1e-6, (l, d, gs)
