"""The AR-RFF part of StandardRoIHead: ``_bbox_forward`` with the 3-region
block enabled (mmdet/models/roi_heads/standard_roi_head.py:135-170; in the
shipped tree the block sits inside a string literal and the head call expects
a 3-tuple -- SURVEY.md section 0 -- the enabled 2-tuple form is what the AR-RFF
config needs and what is provided here).  Assignment, sampling, losses, mask
branch and test-time NMS are orchestration outside the hot path.
"""
import torch.nn as nn

from .bbox_head import MultiBBoxHead, MultiRoIsBBoxHead
from .roi_extractor import SingleRoIExtractor

_EXTRACTORS = {'SingleRoIExtractor': SingleRoIExtractor}
_HEADS = {'MultiBBoxHead': MultiBBoxHead, 'MultiRoIsBBoxHead': MultiRoIsBBoxHead}


def _build(cfg, table):
    if isinstance(cfg, nn.Module):
        return cfg
    cfg = dict(cfg)
    return table[cfg.pop('type')](**cfg)


class StandardRoIHead(nn.Module):

    def __init__(self, bbox_roi_extractor=None, bbox_head=None, facs=1, **unused):
        super(StandardRoIHead, self).__init__()
        self.bbox_roi_extractor = _build(bbox_roi_extractor, _EXTRACTORS)
        self.bbox_head = _build(bbox_head, _HEADS)
        self.facs = facs

    @property
    def with_shared_head(self):
        return False

    def init_weights(self, pretrained=None):
        self.bbox_roi_extractor.init_weights()
        self.bbox_head.init_weights()

    def _bbox_forward(self, x, rois):
        ext = self.bbox_roi_extractor
        bbox_feats = ext.forward_regions(x[:ext.num_inputs], rois, regions=3,
                                         facs=getattr(self, 'facs', 1))
        if getattr(self, 'with_shared_head', False):
            bbox_feats = self.shared_head(bbox_feats)
        cls_score, bbox_pred = self.bbox_head(bbox_feats)[:2]
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)


class CascadeRoIHead(nn.Module):
    """Per-stage ``_bbox_forward(stage, x, rois)`` of the cascade
    (mmdet/models/roi_heads/cascade_roi_head.py:120-142, same 3-region block
    per stage); box refinement between stages is the caller's."""

    def __init__(self, num_stages, bbox_roi_extractor, bbox_head, facs=1, **unused):
        super(CascadeRoIHead, self).__init__()
        def per_stage(c):
            return c if isinstance(c, (list, tuple)) else [c] * num_stages
        self.num_stages = num_stages
        self.bbox_roi_extractor = nn.ModuleList(
            [_build(c, _EXTRACTORS) for c in per_stage(bbox_roi_extractor)])
        self.bbox_head = nn.ModuleList(
            [_build(c, _HEADS) for c in per_stage(bbox_head)])
        self.facs = facs

    def _bbox_forward(self, stage, x, rois):
        ext, head = self.bbox_roi_extractor[stage], self.bbox_head[stage]
        bbox_feats = ext.forward_regions(x[:ext.num_inputs], rois, regions=3,
                                         facs=getattr(self, 'facs', 1))
        cls_score, bbox_pred = head(bbox_feats)[:2]
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)
