"""Region generators of AR-RFF.

``get_adaptive_scale_rois`` keeps the reference's name, arguments and return
order (mmdet/models/utils/additional.py:38-71: returns (adaptive_h, adaptive_w),
which the caller names lh_rois, lw_rois).  The fused extractor evaluates the
same expressions inside the kernel (csrc/geometry.cuh); this host/torch form
exists for API parity, for the unfused composition path and for the tests.
"""
import torch


def get_adaptive_scale_rois(rois, facs):
    b = rois[:, 0:1]
    ctr_x = (rois[:, 1:2] + rois[:, 3:4]) * 0.5
    ctr_y = (rois[:, 2:3] + rois[:, 4:5]) * 0.5
    rw = rois[:, 3:4] - rois[:, 1:2] + 1.0
    rh = rois[:, 4:5] - rois[:, 2:3] + 1.0
    lower = torch.full_like(rw, 0.1)
    half_lh = rh * ((rw / rh) * facs + 1.0) * 0.5
    half_lw = rw * ((rh / rw) * facs + 1.0) * 0.5
    half_w = rw * 0.5
    adaptive_h = torch.cat((b, torch.max(ctr_x - half_w, lower),
                            torch.max(ctr_y - half_lh, lower),
                            ctr_x + half_w, ctr_y + half_lh), dim=-1)
    adaptive_w = torch.cat((b, torch.max(ctr_x - half_lw, lower),
                            torch.max(ctr_y - half_lh, lower),
                            ctr_x + half_lw, ctr_y + half_lh), dim=-1)
    return adaptive_h, adaptive_w
