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


def accelerate_class(ref_cls, ours, methods):
    """Subclass of a reference class whose hot-path methods are replaced by ours and
    everything else (losses, targets, box decoding, config handling, parameter
    construction) stays the reference's.  Our methods only touch attributes the
    reference classes define under the same names (hh_conv / wh_conv / final_conv /
    shared_fcs ..., reduce_convs / reduce_convs2 / refine, bbox_roi_extractor /
    bbox_head), so checkpoints and configs are unaffected."""
    if getattr(ref_cls, '_arfe_b200_base', None) is not None:
        return ref_cls  # already accelerated
    body = {m: ours.__dict__[m] for m in methods}
    body['_arfe_b200_base'] = ref_cls
    body['__module__'] = ours.__module__
    body['__doc__'] = f'{ref_cls.__name__} with {", ".join(methods)} from arfe_b200.'
    return type(ref_cls.__name__, (ref_cls,), body)


def register_into_mmdet(force=True):
    """Swap the B200 path into mmdet's registries under the reference's names, so the
    released configs build it unchanged (mmdet/models/builder.py:4-10).

    * necks / bbox heads / RoI heads that exist in the registry are SUBCLASSED
      (accelerate_class): WFPNDualSpatial.forward, MultiBBoxHead.forward/.fuse and
      Standard/CascadeRoIHead._bbox_forward become ours; loss(), get_targets(),
      get_bboxes(), refine_bboxes, bbox_coder, train/test orchestration remain the
      reference's.  (Registering our stand-alone heads instead would build, then fail
      at the first loss() call.)
    * SingleRoIExtractor and the RoIAlign op have no trainable state and no loss
      code; ours are registered as they are.
    Returns False when mmdet is not installed."""
    try:
        from mmdet import ops as mm_ops
        from mmdet.models.builder import HEADS, NECKS, ROI_EXTRACTORS
    except Exception:
        return False
    from . import (CascadeRoIHead, MultiBBoxHead, RoIAlign, SingleRoIExtractor,
                   StandardRoIHead, WFPNDualSpatial)

    def put(reg, cls):
        mods = getattr(reg, '_module_dict', None)
        if mods is not None and force:
            mods[cls.__name__] = cls
        else:
            reg.register_module(cls)

    def swap(reg, name, ours, methods):
        mods = getattr(reg, '_module_dict', {})
        ref_cls = mods.get(name)
        if ref_cls is None:
            if name == ours.__name__:
                put(reg, ours)
            return
        put(reg, accelerate_class(ref_cls, ours, methods))

    swap(NECKS, 'WFPNDualSpatial', WFPNDualSpatial, ('forward',))
    for name in ('MultiBBoxHead', 'MultiRoIsBBoxHead'):
        swap(HEADS, name, MultiBBoxHead, ('forward', 'fuse'))
    swap(HEADS, 'StandardRoIHead', StandardRoIHead, ('_bbox_forward',))
    swap(HEADS, 'CascadeRoIHead', CascadeRoIHead, ('_bbox_forward',))
    put(ROI_EXTRACTORS, SingleRoIExtractor)
    mm_ops.RoIAlign = RoIAlign
    return True


def optimize_detector(model, fused_attention='auto'):
    """Put a built detector on the fast path: parameters and activations in
    torch.channels_last (cuDNN's preferred layout for the convs around the path, and
    the layout in which our kernels need no transposes and no atomics) and the RoI
    extractors in split mode (regions handed to the head as separate tensors).
    Without it an unchanged NCHW config runs through the compatibility kernels --
    correct, ~3.8x slower on the RoI part.
    fused_attention: NonLocal2D.fused_attention of every refine block ('auto': the tensor-core
    attention for bfloat16 activations only; True: also for float32 ones, whose operands are
    then rounded to bf16)."""
    import torch
    from .neck import NonLocal2D
    from .roi_extractor import SingleRoIExtractor
    model = model.to(memory_format=torch.channels_last)
    for m in model.modules():
        if isinstance(m, SingleRoIExtractor):
            m.roi_feats_split = True
            m.roi_feats_channels_last = True
        elif isinstance(m, NonLocal2D):
            m.fused_attention = fused_attention
    return model
