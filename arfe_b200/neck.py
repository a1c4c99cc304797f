"""AR-FPN neck with the reference's surface and parameter names
(mmdet/models/necks/wfpn_dual_spatial.py:10-137): reduce_convs.{i}.conv,
reduce_convs2.{i}.conv, refine.{g,theta,phi,conv_out}.conv.

The gather (all levels -> refine level, mean) and the gated residual
``x + up(bsf) * (tanh(relu(c1)) + tanh(relu(c2)))`` run in libarfe_b200.so;
the NonLocal2D refine's attention is one tensor-core kernel (tcgen05) when its operands may be
bf16, and both C -> 1 gate convolutions of every level one pass over the pyramid (forward); the
other convolutions are library calls.
"""
import torch
import torch.nn as nn

from ._compat import ConvModule, xavier_init
from . import _lib as L
from .functional import FPNLink, fpn_apply, fpn_gate_conv, fpn_gather, nonlocal_attention


class NonLocal2D(nn.Module):
    """mmdet/ops/non_local.py:6-105 (embedded gaussian / dot product), same constructor and
    parameter names (g / theta / phi / conv_out).

    ``fused_attention`` selects how softmax(theta_x . phi_x) . g_x (:65-69, :98-101) runs:
      True    one tensor-core attention kernel (arfe_nonlocal_attention_forward): operands rounded
              to bf16, fp32 accumulation and softmax; the HW x HW weight matrix is never materialised;
      False   the reference's two matmuls and softmax through the library;
      'auto'  (default) fused for bfloat16 activations, library for float32 -- an fp32 module keeps
              the reference's fp32 arithmetic unless the caller opts into bf16 operands
              (``arfe_b200.optimize_detector(model, fused_attention=True)``), or has already told
              PyTorch that bf16 is acceptable inside float32 matmuls
              (``torch.set_float32_matmul_precision('medium')``).
    The 1x1 convolutions stay library GEMMs either way."""

    def __init__(self, in_channels, reduction=2, use_scale=True, conv_cfg=None,
                 norm_cfg=None, mode='embedded_gaussian'):
        super(NonLocal2D, self).__init__()
        assert mode in ['embedded_gaussian', 'dot_product']
        self.in_channels, self.reduction, self.use_scale = in_channels, reduction, use_scale
        self.inter_channels = in_channels // reduction
        self.mode = mode
        self.g = ConvModule(in_channels, self.inter_channels, kernel_size=1, act_cfg=None)
        self.theta = ConvModule(in_channels, self.inter_channels, kernel_size=1, act_cfg=None)
        self.phi = ConvModule(in_channels, self.inter_channels, kernel_size=1, act_cfg=None)
        self.conv_out = ConvModule(self.inter_channels, in_channels, kernel_size=1,
                                   conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=None)
        self.fused_attention = 'auto'
        self.init_weights()

    def _use_fused(self, x):
        if self.fused_attention is False or self.mode != 'embedded_gaussian':
            return False
        ok = x.is_cuda and self.inter_channels in (64, 128, 256) and \
            x.dtype in (torch.float32, torch.bfloat16)
        if self.fused_attention is True:
            if not ok:
                raise RuntimeError("NonLocal2D(fused_attention=True) needs CUDA float32/bfloat16 "
                                   "activations and inter_channels in (64, 128, 256)")
            return True
        return ok and (x.dtype == torch.bfloat16 or torch.get_float32_matmul_precision() == 'medium')

    def init_weights(self, std=0.01, zeros_init=True):
        for m in [self.g, self.theta, self.phi]:
            nn.init.normal_(m.conv.weight, 0, std)
            nn.init.constant_(m.conv.bias, 0)
        if zeros_init:
            nn.init.constant_(self.conv_out.conv.weight, 0)
        else:
            nn.init.normal_(self.conv_out.conv.weight, 0, std)
        nn.init.constant_(self.conv_out.conv.bias, 0)

    def forward(self, x):
        n, _, h, w = x.shape
        if self._use_fused(x):
            scale = 1.0 / self.inter_channels ** 0.5 if self.use_scale else 1.0
            y = nonlocal_attention(self.theta(x), self.phi(x), self.g(x), scale)
            return x + self.conv_out(y)
        g_x = self.g(x).reshape(n, self.inter_channels, -1).permute(0, 2, 1)
        theta_x = self.theta(x).reshape(n, self.inter_channels, -1).permute(0, 2, 1)
        phi_x = self.phi(x).reshape(n, self.inter_channels, -1)
        pw = torch.matmul(theta_x, phi_x)
        if self.mode == 'embedded_gaussian':
            if self.use_scale:
                pw = pw / theta_x.shape[-1] ** 0.5
            pw = pw.softmax(dim=-1)
        else:
            pw = pw / pw.shape[-1]
        y = torch.matmul(pw, g_x).permute(0, 2, 1).contiguous().reshape(
            n, self.inter_channels, h, w)
        return x + self.conv_out(y)


class WFPNDualSpatial(nn.Module):

    def __init__(self, in_channels, num_levels, refine_level=2, conv_cfg=None,
                 norm_cfg=None):
        super(WFPNDualSpatial, self).__init__()
        self.in_channels = in_channels
        self.num_levels = num_levels
        self.conv_cfg = conv_cfg
        self.norm_cfg = norm_cfg
        self.refine_level = refine_level
        self.reduce_convs = nn.ModuleList()
        self.reduce_convs2 = nn.ModuleList()
        for _ in range(num_levels):
            for convs in (self.reduce_convs, self.reduce_convs2):
                convs.append(ConvModule(in_channels, 1, 3, padding=1, conv_cfg=conv_cfg,
                                        norm_cfg=norm_cfg, inplace=False))
        self.refine = NonLocal2D(in_channels, reduction=1, use_scale=False,
                                 conv_cfg=conv_cfg, norm_cfg=norm_cfg)
        # channels-last inputs: the 2 x num_levels gate convolutions as one fused pass (forward);
        # set False to run them through cuDNN like the reference
        self.fuse_gate_convs = True

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution='uniform')

    def _gate_convs_fusable(self, inputs):
        x = inputs[0]
        v = 4 if x.dtype == torch.float32 else 8
        return x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 4 and \
            x.shape[1] % v == 0 and x.shape[1] <= 64 * v and \
            all(L.layout_of(t) == L.ARFE_NHWC for t in inputs) and \
            all(tuple(m.conv.weight.shape[2:]) == (3, 3) and m.conv.bias is not None
                for m in list(self.reduce_convs) + list(self.reduce_convs2))

    def forward(self, inputs):
        assert len(inputs) == self.num_levels
        # bsf is computed from the gather's output, so the two ops can share one write of d x_l
        link = FPNLink() if torch.is_grad_enabled() else None
        ori_fe = fpn_gather(inputs, self.refine_level, link)
        bsf = self.refine(ori_fe)
        # pre-activation gate maps; relu (ConvModule's default act) + tanh + sum
        # are fused into the apply kernel
        if self.fuse_gate_convs and self._gate_convs_fusable(inputs):
            # both C -> 1 convolutions of every level in one pass over the pyramid
            c1 = [m.conv for m in self.reduce_convs]
            c2 = [m.conv for m in self.reduce_convs2]
            g1, g2 = fpn_gate_conv(list(inputs), [c.weight for c in c1], [c.bias for c in c1],
                                   [c.weight for c in c2], [c.bias for c in c2])
        else:
            g1 = [self.reduce_convs[i](inputs[i], activate=False) for i in range(self.num_levels)]
            g2 = [self.reduce_convs2[i](inputs[i], activate=False) for i in range(self.num_levels)]
        return tuple(fpn_apply(list(inputs), bsf, g1, g2, link))
