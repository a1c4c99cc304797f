"""AR-RFF fusion head with the reference's surface and parameter names
(mmdet/models/roi_heads/bbox_heads/multirois_bbox_head.py:13-251):
hh_conv/wh_conv/final_conv (.conv.weight/.bias), shared_fcs.{i}, fc_cls,
fc_reg.  The 3x3 convolutions and FCs stay on cuDNN/cuBLAS (north_star); the
channel split and the gate ``ori*(1+a+b)`` run in libarfe_b200.so.
Loss/target/decoding code of BBoxHead (bbox_head.py:91-352) is orchestration
outside the hot path and is not reproduced.
"""
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair

from ._compat import ConvModule, xavier_init
from .functional import rff_gate, split3


class MultiBBoxHead(nn.Module):

    def __init__(self, num_shared_convs=0, num_shared_fcs=0, num_cls_convs=0,
                 num_cls_fcs=0, num_reg_convs=0, num_reg_fcs=0,
                 conv_out_channels=256, fc_out_channels=1024, num_ws_convs=2,
                 num_ws_fcs=2, conv_cfg=None, norm_cfg=None,
                 with_avg_pool=False, with_cls=True, with_reg=True,
                 roi_feat_size=7, in_channels=256, num_classes=80,
                 reg_class_agnostic=False, **unused_loss_and_coder_cfg):
        super(MultiBBoxHead, self).__init__()
        assert (num_shared_convs + num_shared_fcs + num_cls_convs +
                num_cls_fcs + num_reg_convs + num_reg_fcs > 0)
        assert with_cls or with_reg
        self.with_avg_pool, self.with_cls, self.with_reg = with_avg_pool, with_cls, with_reg
        self.roi_feat_size = _pair(roi_feat_size)
        self.roi_feat_area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.in_channels, self.num_classes = in_channels, num_classes
        self.reg_class_agnostic = reg_class_agnostic
        self.fp16_enabled = False
        self.num_shared_convs, self.num_shared_fcs = num_shared_convs, num_shared_fcs
        self.num_cls_convs, self.num_cls_fcs = num_cls_convs, num_cls_fcs
        self.num_reg_convs, self.num_reg_fcs = num_reg_convs, num_reg_fcs
        self.conv_out_channels, self.fc_out_channels = conv_out_channels, fc_out_channels
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        if with_avg_pool:
            self.avg_pool = nn.AvgPool2d(self.roi_feat_size)

        def conv3():
            return ConvModule(in_channels, in_channels, 3, padding=1,
                              conv_cfg=conv_cfg, norm_cfg=norm_cfg, inplace=False)
        self.hh_conv, self.wh_conv, self.final_conv = conv3(), conv3(), conv3()

        self.shared_convs, self.shared_fcs, last = self._add_conv_fc_branch(
            num_shared_convs, num_shared_fcs, in_channels, True)
        self.shared_out_channels = last
        self.cls_convs, self.cls_fcs, self.cls_last_dim = self._add_conv_fc_branch(
            num_cls_convs, num_cls_fcs, last)
        self.reg_convs, self.reg_fcs, self.reg_last_dim = self._add_conv_fc_branch(
            num_reg_convs, num_reg_fcs, last)
        if num_shared_fcs == 0 and not with_avg_pool:
            if num_cls_fcs == 0:
                self.cls_last_dim *= self.roi_feat_area
            if num_reg_fcs == 0:
                self.reg_last_dim *= self.roi_feat_area
        self.relu = nn.ReLU(inplace=True)
        if with_cls:
            self.fc_cls = nn.Linear(self.cls_last_dim, num_classes + 1)
        if with_reg:
            self.fc_reg = nn.Linear(
                self.reg_last_dim, 4 if reg_class_agnostic else 4 * num_classes)

    def _add_conv_fc_branch(self, num_convs, num_fcs, in_channels, is_shared=False):
        last = in_channels
        convs = nn.ModuleList()
        for i in range(num_convs):
            convs.append(ConvModule(last if i == 0 else self.conv_out_channels,
                                    self.conv_out_channels, 3, padding=1,
                                    conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
        if num_convs > 0:
            last = self.conv_out_channels
        fcs = nn.ModuleList()
        if num_fcs > 0:
            if (is_shared or self.num_shared_fcs == 0) and not self.with_avg_pool:
                last *= self.roi_feat_area
            for i in range(num_fcs):
                fcs.append(nn.Linear(last if i == 0 else self.fc_out_channels,
                                     self.fc_out_channels))
            last = self.fc_out_channels
        return convs, fcs, last

    def init_weights(self):
        if self.with_cls:
            nn.init.normal_(self.fc_cls.weight, 0, 0.01)
            nn.init.constant_(self.fc_cls.bias, 0)
        if self.with_reg:
            nn.init.normal_(self.fc_reg.weight, 0, 0.001)
            nn.init.constant_(self.fc_reg.bias, 0)
        for fcs in (self.shared_fcs, self.cls_fcs, self.reg_fcs):
            for m in fcs.modules():
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight)
                    nn.init.constant_(m.bias, 0)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution='uniform')

    def fuse(self, x):
        """[K, 3C, h, w] (or the three region tensors (ori, lw, lh) of the
        extractor's split mode) -> gated [K, C, h, w] (multirois_bbox_head.py:167-182)."""
        ori, lwh, lhh = x if isinstance(x, (tuple, list)) else split3(x, self.conv_out_channels)
        a = F.relu(self.wh_conv(lwh))
        b = F.relu(self.hh_conv(lhh))
        return rff_gate(ori, a, b)

    def forward(self, x):
        x_out = F.relu(self.final_conv(self.fuse(x)))
        for conv in self.shared_convs:
            x_out = conv(x_out)
        if self.num_shared_fcs > 0:
            if self.with_avg_pool:
                x_out = self.avg_pool(x_out)
            x_out = x_out.flatten(1)
            for fc in self.shared_fcs:
                x_out = self.relu(fc(x_out))
        x_cls = x_reg = x_out
        for conv in self.cls_convs:
            x_cls = conv(x_cls)
        if x_cls.dim() > 2:
            x_cls = (self.avg_pool(x_cls) if self.with_avg_pool else x_cls).flatten(1)
        for fc in self.cls_fcs:
            x_cls = self.relu(fc(x_cls))
        for conv in self.reg_convs:
            x_reg = conv(x_reg)
        if x_reg.dim() > 2:
            x_reg = (self.avg_pool(x_reg) if self.with_avg_pool else x_reg).flatten(1)
        for fc in self.reg_fcs:
            x_reg = self.relu(fc(x_reg))
        cls_score = self.fc_cls(x_cls) if self.with_cls else None
        bbox_pred = self.fc_reg(x_reg) if self.with_reg else None
        return cls_score, bbox_pred


class MultiRoIsBBoxHead(MultiBBoxHead):
    """Preset of the released config (multirois_bbox_head.py:238-251)."""

    def __init__(self, fc_out_channels=1024, *args, **kwargs):
        super(MultiRoIsBBoxHead, self).__init__(
            num_shared_convs=0, num_shared_fcs=2, num_cls_convs=0,
            num_cls_fcs=0, num_reg_convs=0, num_reg_fcs=0,
            fc_out_channels=fc_out_channels, *args, **kwargs)
