"""The little of mmcv the path relies on, for environments without mmcv/mmdet
(SURVEY.md section 8(c)): ``ConvModule`` = Conv2d(bias=True) followed by ReLU
unless ``act_cfg=None``, parameters under ``.conv`` so reference checkpoints
load unchanged; ``xavier_init``; and registration into mmdet's registries when
mmdet is importable (mmdet/models/builder.py:4-10).
"""
import torch.nn as nn
import torch.nn.functional as F


class ConvModule(nn.Module):
    _DEFAULT_ACT = dict(type='ReLU')

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias='auto', conv_cfg=None,
                 norm_cfg=None, act_cfg=_DEFAULT_ACT, inplace=True):
        super().__init__()
        if conv_cfg is not None or norm_cfg is not None:
            raise NotImplementedError(
                "conv_cfg/norm_cfg need mmcv's ConvModule; the ARFE configs "
                "leave both None (wfpn_dual_spatial.py:19-20)")
        assert act_cfg is None or act_cfg.get('type') == 'ReLU'
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride,
                              padding, dilation, groups,
                              bias=True if bias == 'auto' else bias)
        self.with_activation = act_cfg is not None

    def forward(self, x, activate=True):
        x = self.conv(x)
        if activate and self.with_activation:
            x = F.relu(x)
        return x


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def register_into_mmdet(force=True):
    """Swap the B200 modules into mmdet's registries under the reference's
    names, so existing configs (type='WFPNDualSpatial', 'SingleRoIExtractor',
    'MultiRoIsBBoxHead', roi_layer type 'RoIAlign') build them unchanged.
    Returns False when mmdet is not installed."""
    try:
        from mmdet import ops as mm_ops
        from mmdet.models.builder import HEADS, NECKS, ROI_EXTRACTORS
    except Exception:
        return False
    from . import (MultiBBoxHead, MultiRoIsBBoxHead, RoIAlign,
                   SingleRoIExtractor, WFPNDualSpatial)
    for reg, cls in ((NECKS, WFPNDualSpatial), (ROI_EXTRACTORS, SingleRoIExtractor),
                     (HEADS, MultiBBoxHead), (HEADS, MultiRoIsBBoxHead)):
        mods = getattr(reg, '_module_dict', None)
        if mods is not None and force:
            mods[cls.__name__] = cls
        else:
            reg.register_module(cls)
    mm_ops.RoIAlign = RoIAlign
    return True
