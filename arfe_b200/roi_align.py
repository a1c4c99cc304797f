"""RoIAlign operator with the reference's module/function surface.

Mirrors mmdet/ops/roi_align/roi_align.py:76-154 (``roi_align`` functional and
``RoIAlign`` module: same constructor arguments, attributes and repr); the
compute is arfe_roi_align_forward/backward in libarfe_b200.so.
"""
from torch import nn
from torch.nn.modules.utils import _pair

from .functional import RoIAlignFunction, roi_align  # noqa: F401


class RoIAlign(nn.Module):

    def __init__(self, out_size, spatial_scale, sample_num=0,
                 use_torchvision=False, aligned=True):
        super(RoIAlign, self).__init__()
        self.out_size = _pair(out_size)
        self.spatial_scale = float(spatial_scale)
        self.aligned = aligned
        self.sample_num = int(sample_num)
        # kept as an attribute for config / repr compatibility (roi_align.py:110-124 of the
        # reference); there is one implementation here, so asking for another one is an error
        self.use_torchvision = use_torchvision
        if use_torchvision:
            raise NotImplementedError(
                'use_torchvision=True selects torchvision.ops.roi_align in the reference; '
                'arfe_b200 has a single implementation (libarfe_b200.so) and no library fallback')

    def forward(self, features, rois):
        """features: NCHW (or channels_last) map; rois: [K,5] (idx,x1,y1,x2,y2)."""
        assert rois.dim() == 2 and rois.size(1) == 5
        return roi_align(features, rois, self.out_size, self.spatial_scale,
                         self.sample_num, self.aligned)

    def __repr__(self):
        indent_str = '\n    '
        format_str = self.__class__.__name__
        format_str += f'({indent_str}out_size={self.out_size},'
        format_str += f'{indent_str}spatial_scale={self.spatial_scale},'
        format_str += f'{indent_str}sample_num={self.sample_num},'
        format_str += f'{indent_str}use_torchvision={self.use_torchvision},'
        format_str += f'{indent_str}aligned={self.aligned})'
        return format_str
