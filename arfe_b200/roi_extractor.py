"""SingleRoIExtractor with the reference's constructor/forward surface
(mmdet/models/roi_heads/roi_extractors/single_level.py:10-152), backed by the
fused multi-level kernel, plus ``forward_regions`` -- the whole AR-RFF
extraction (3 region boxes x level map x RoIAlign x cat) in one launch.
"""
import torch
import torch.nn as nn

from .functional import roi_fuse, roi_fuse_split
from .roi_align import RoIAlign

_ROI_LAYERS = {'RoIAlign': RoIAlign}


class SingleRoIExtractor(nn.Module):
    """Args as in the reference: roi_layer (dict with ``type`` looked up by
    name, single_level.py:44-51), out_channels, featmap_strides, finest_scale."""

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super(SingleRoIExtractor, self).__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale
        self.fp16_enabled = False
        # emit RoI features in torch.channels_last when the pyramid is channels-last
        self.roi_feats_channels_last = False
        # forward_regions returns (ori, lw, lh) instead of their cat for channels-last pyramids
        # (opt-in: the reference interface is the concatenated tensor)
        self.roi_feats_split = False

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        assert layer_type in _ROI_LAYERS, \
            f'roi_layer type {layer_type!r} is not provided (have {list(_ROI_LAYERS)})'
        layer_cls = _ROI_LAYERS[layer_type]
        return nn.ModuleList(
            [layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def map_roi_levels(self, rois, num_levels):
        """single_level.py:53-93 in torch ops (used only by the hook path; the
        fused kernel evaluates the same rule in csrc/geometry.cuh)."""
        scale = torch.sqrt(
            (rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def roi_rescale(self, rois, scale_factor):
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        new_w = (rois[:, 3] - rois[:, 1]) * scale_factor
        new_h = (rois[:, 4] - rois[:, 2]) * scale_factor
        return torch.stack((rois[:, 0], cx - new_w * 0.5, cy - new_h * 0.5,
                            cx + new_w * 0.5, cy + new_h * 0.5), dim=-1)

    def _layer_args(self):
        l0 = self.roi_layers[0]
        if not getattr(l0, 'aligned', True) or getattr(l0, 'use_torchvision', False):
            raise NotImplementedError('fused extractor needs aligned RoIAlign layers')
        return l0.out_size, l0.sample_num, [l.spatial_scale for l in self.roi_layers]

    def _cast_in(self, feats):
        # @force_fp32(apply_to=('feats',), out_fp16=True), single_level.py:109
        half = any(f.dtype == torch.float16 for f in feats)
        if half:
            feats = [f.float() for f in feats]
        return list(feats), half

    def forward(self, feats, rois, roi_scale_factor=None, lvl=None,
                replace_rois=None):
        out_size, sample_num, scales = self._layer_args()
        feats, half = self._cast_in(feats)
        num_levels = len(feats)
        if roi_scale_factor is None and lvl is None and replace_rois is None:
            out = roi_fuse(feats, rois, out_size, scales[:num_levels], sample_num,
                           regions=1, finest_scale=self.finest_scale,
                           out_channels_last=self.roi_feats_channels_last)
        else:
            # hook path (unused by ARFE's configs): the reference's order of operations,
            # single_level.py:118-152 -- levels from the ORIGINAL rois (or replace_rois),
            # shifted by lvl, and only then the rescale of the boxes that are sampled;
            # one single-level launch of the same kernel per level
            out = feats[0].new_zeros(rois.size(0), self.out_channels, *out_size)
            if num_levels == 1:  # :121-124 returns before the rescale
                out = self.roi_layers[0](feats[0], rois) if len(rois) else out
            else:
                src = replace_rois if replace_rois is not None else rois
                target = self.map_roi_levels(src, num_levels)
                if lvl is not None:
                    target = (target + lvl).clamp(min=0, max=num_levels - 1).long()
                if roi_scale_factor is not None:
                    rois = self.roi_rescale(rois, roi_scale_factor)
                for i in range(num_levels):
                    inds = target == i
                    if inds.any():
                        out[inds] = self.roi_layers[i](feats[i], rois[inds, :])
        return out.half() if half else out

    def forward_regions(self, feats, rois, regions=3, facs=1.0):
        """AR-RFF extraction: cat([ori, lw, lh], dim=1) of the three region
        boxes' features (standard_roi_head.py:138-155), one kernel launch."""
        out_size, sample_num, scales = self._layer_args()
        feats, half = self._cast_in(feats)
        if self.roi_feats_split and all(f.is_contiguous(memory_format=torch.channels_last) for f in feats):
            # the regions as separate tensors: the head convolves each on its own, so the
            # cat (forward) and the slice gradients (backward) would be pure copies
            outs = roi_fuse_split(feats, rois, out_size, scales[:len(feats)], sample_num,
                                  regions=regions, facs=facs, finest_scale=self.finest_scale)
            return tuple(o.half() for o in outs) if half else outs
        out = roi_fuse(feats, rois, out_size, scales[:len(feats)], sample_num,
                       regions=regions, facs=facs, finest_scale=self.finest_scale,
                       out_channels_last=self.roi_feats_channels_last)
        return out.half() if half else out
