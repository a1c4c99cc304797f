"""oracle/arfe_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's region-aware feature path, used only as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under arfe_b200/ imports this module.

The Python parts of the reference cannot be imported here (mmdet needs mmcv,
which is un-vendored and not installed), so they are
restated with the very torch CPU ops the reference calls -- the arithmetic is
executed by the same ATen kernels, not re-derived.  RoIAlign itself runs
either through the reference's own C++ compiled unmodified (``oracle/_ref``,
backend "ref") or through our plain-C restatement (``roi_align_oracle.c``,
backend "c"); tests pin the two (and torchvision) bit-for-bit.

Pinning status: the reference ships NO golden vectors or tests for this path
(SURVEY.md section 4).  The oracle is pinned against (i) the reference's own
RoIAlign compiled here, (ii) torchvision.ops.roi_align, and (iii) fixtures in
tests/golden/ generated from (i) by tests/golden/make_golden.py.

File:line citations are relative to /root/reference/.
"""
import ctypes
import importlib.util
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------
# RoIAlign back-ends
# --------------------------------------------------------------------------
_c_lib = None
_ref_ext = None


def c_lib():
    """ctypes handle of liboracle_roialign.so (built by build_oracle.py)."""
    global _c_lib
    if _c_lib is None:
        path = os.path.join(_HERE, "liboracle_roialign.so")
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build_c_oracle()
        lib = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.oracle_roi_align_forward.argtypes = [
            fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp,
            ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, fp]
        lib.oracle_roi_align_backward.argtypes = [
            fp, fp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, fp]
        lib.oracle_roi_align_taps.argtypes = [
            fp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ip, ip, ip, fp, fp, ip, ip, fp, fp]
        _c_lib = lib
    return _c_lib


def ref_ext():
    """The reference's roi_align_ext compiled unmodified, or None."""
    global _ref_ext
    if _ref_ext is None:
        from . import build_oracle
        path = build_oracle.ref_ext_path()
        if path is None:
            return None
        spec = importlib.util.spec_from_file_location("roi_align_ext", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_ext = mod
    return _ref_ext


_ref_nms = None


def ref_nms_ext():
    """The reference's nms_ext (CPU build) compiled unmodified, or None."""
    global _ref_nms
    if _ref_nms is None:
        from . import build_oracle
        path = build_oracle.ref_nms_ext_path()
        if path is None:
            return None
        spec = importlib.util.spec_from_file_location("nms_ext", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_nms = mod
    return _ref_nms


def nms(dets, iou_thr, backend="py"):
    """ops/nms/src/cpu/nms_cpu.cpp:8-69 (greedy NMS in descending-score order, IoU with
    areas (x2-x1)*(y2-y1), suppressed when ovr > threshold): kept indices into dets, in
    score order.  backend "ref" runs the reference's own compiled code; "py" restates it
    with numpy float32 scalars (each product / sum rounded separately, as the C++ does)."""
    dets = dets.detach().float().cpu().contiguous()
    if backend == "ref":
        return ref_nms_ext().nms(dets, float(iou_thr))
    d = dets.numpy().astype(np.float32)
    if d.shape[0] == 0:
        return torch.zeros(0, dtype=torch.long)
    x1, y1, x2, y2, sc = (d[:, i] for i in range(5))
    areas = (x2 - x1) * (y2 - y1)
    order = torch.from_numpy(sc).sort(0, descending=True)[1].numpy()
    suppressed = np.zeros(d.shape[0], dtype=bool)
    keep = []
    thr = np.float32(iou_thr)
    for _i in range(d.shape[0]):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(int(i))
        rest = order[_i + 1:]
        xx1, yy1 = np.maximum(x1[i], x1[rest]), np.maximum(y1[i], y1[rest])
        xx2, yy2 = np.minimum(x2[i], x2[rest]), np.minimum(y2[i], y2[rest])
        w, h = np.maximum(np.float32(0), xx2 - xx1), np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return torch.tensor(keep, dtype=torch.long)


def bbox2roi(bbox_list):
    """core/bbox/transforms.py:41-60."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def _fptr(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _iptr(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int32))


def roi_align_forward(feat, rois, out_size, spatial_scale, sample_num=0,
                      backend="c"):
    """roi_align_ext.forward_v2 (ops/roi_align/src/roi_align_ext.cpp:126-142)."""
    oh, ow = (out_size, out_size) if isinstance(out_size, int) else out_size
    feat = feat.detach().float().contiguous()
    rois = rois.detach().float().contiguous()
    if backend == "ref":
        return ref_ext().forward_v2(feat, rois, float(spatial_scale), oh, ow,
                                    int(sample_num), True)
    B, C, H, W = feat.shape
    out = torch.zeros(rois.shape[0], C, oh, ow)
    if out.numel():
        rc = c_lib().oracle_roi_align_forward(
            _fptr(feat), B, C, H, W, _fptr(rois), rois.shape[0],
            float(spatial_scale), oh, ow, int(sample_num), _fptr(out))
        if rc != 0:
            raise RuntimeError("ROIs in ROIAlign cannot have negative size")
    return out


def roi_align_backward(dout, rois, out_size, spatial_scale, feat_shape,
                       sample_num=0, backend="c"):
    """roi_align_ext.backward_v2 (roi_align_ext.cpp:144-161)."""
    oh, ow = (out_size, out_size) if isinstance(out_size, int) else out_size
    B, C, H, W = feat_shape
    dout = dout.detach().float().contiguous()
    rois = rois.detach().float().contiguous()
    if backend == "ref":
        return ref_ext().backward_v2(dout, rois, float(spatial_scale), oh, ow,
                                     B, C, H, W, int(sample_num), True)
    din = torch.zeros(B, C, H, W)
    if dout.numel():
        rc = c_lib().oracle_roi_align_backward(
            _fptr(dout), _fptr(rois), rois.shape[0], float(spatial_scale), oh,
            ow, B, C, H, W, int(sample_num), _fptr(din))
        if rc != 0:
            raise RuntimeError("ROIs in ROIAlign cannot have negative size")
    return din


def roi_align_taps(rois, out_size, spatial_scale, H, W, sample_num=0,
                   max_grid=16):
    """Sampling rows/columns/weights of every RoI (for the index parity test)."""
    oh, ow = (out_size, out_size) if isinstance(out_size, int) else out_size
    rois = rois.detach().float().contiguous()
    K = rois.shape[0]
    grid = torch.zeros(K, 2, dtype=torch.int32)
    ylo = torch.zeros(K, oh, max_grid, dtype=torch.int32)
    yhi = torch.zeros_like(ylo)
    ywl = torch.zeros(K, oh, max_grid)
    ywh = torch.zeros_like(ywl)
    xlo = torch.zeros(K, ow, max_grid, dtype=torch.int32)
    xhi = torch.zeros_like(xlo)
    xwl = torch.zeros(K, ow, max_grid)
    xwh = torch.zeros_like(xwl)
    c_lib().oracle_roi_align_taps(
        _fptr(rois), K, float(spatial_scale), oh, ow, H, W, int(sample_num),
        max_grid, _iptr(grid), _iptr(ylo), _iptr(yhi), _fptr(ywl), _fptr(ywh),
        _iptr(xlo), _iptr(xhi), _fptr(xwl), _fptr(xwh))
    return dict(grid=grid, ylo=ylo, yhi=yhi, ywl=ywl, ywh=ywh, xlo=xlo,
                xhi=xhi, xwl=xwl, xwh=xwh)


class _RoIAlignFn(torch.autograd.Function):
    """ops/roi_align/roi_align.py:9-73 (aligned=True branch)."""

    @staticmethod
    def forward(ctx, feat, rois, out_size, spatial_scale, sample_num, backend):
        ctx.save_for_backward(rois)
        ctx.meta = (out_size, spatial_scale, sample_num, tuple(feat.shape),
                    backend)
        return roi_align_forward(feat, rois, out_size, spatial_scale,
                                 sample_num, backend)

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        out_size, scale, sn, shape, backend = ctx.meta
        return (roi_align_backward(g, rois, out_size, scale, shape, sn,
                                   backend), None, None, None, None, None)


def roi_align(feat, rois, out_size, spatial_scale, sample_num=0, backend="c"):
    return _RoIAlignFn.apply(feat, rois, out_size, spatial_scale, sample_num,
                             backend)


# --------------------------------------------------------------------------
# AR-RFF: regions, level map, extractor, assembly, gate
# --------------------------------------------------------------------------
def get_adaptive_scale_rois(rois, facs=1):
    """models/utils/additional.py:38-71, op for op."""
    ctr_x = ((rois[:, 1] + rois[:, 3]) * 0.5).view(-1, 1)
    ctr_y = ((rois[:, 2] + rois[:, 4]) * 0.5).view(-1, 1)
    rw = (rois[:, 3] - rois[:, 1] + 1.0).view(-1, 1)
    rh = (rois[:, 4] - rois[:, 2] + 1.0).view(-1, 1)
    floor_c = torch.ones_like(rw) * 0.1
    h_rate = (rw / rh) * facs + 1.0
    w_rate = (rh / rw) * facs + 1.0
    large_h = rh * h_rate
    large_w = rw * w_rate
    b = rois[:, 0].view(-1, 1)
    adaptive_h = torch.cat((b,
                            torch.max(ctr_x - rw * 0.5, floor_c),
                            torch.max(ctr_y - large_h * 0.5, floor_c),
                            ctr_x + rw * 0.5,
                            ctr_y + large_h * 0.5), dim=-1)
    adaptive_w = torch.cat((b,
                            torch.max(ctr_x - large_w * 0.5, floor_c),
                            torch.max(ctr_y - large_h * 0.5, floor_c),
                            ctr_x + large_w * 0.5,
                            ctr_y + large_h * 0.5), dim=-1)
    return adaptive_h, adaptive_w


def map_roi_levels(rois, num_levels, finest_scale=56):
    """roi_extractors/single_level.py:53-93."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lvls.clamp(min=0, max=num_levels - 1).long()


def single_roi_extractor(feats, rois, featmap_strides, out_size=7,
                         sample_num=0, finest_scale=56, backend="c"):
    """roi_extractors/single_level.py:109-152 (hooks unused by ARFE omitted:
    roi_scale_factor / lvl / replace_rois are all None at every call site)."""
    oh, ow = (out_size, out_size) if isinstance(out_size, int) else out_size
    num_levels = len(feats)
    C = feats[0].shape[1]
    roi_feats = feats[0].new_zeros(rois.size(0), C, oh, ow)
    if num_levels == 1:
        if len(rois) == 0:
            return roi_feats
        return roi_align(feats[0], rois, (oh, ow), 1 / featmap_strides[0],
                         sample_num, backend)
    lvls = map_roi_levels(rois, num_levels, finest_scale)
    for i in range(num_levels):
        inds = lvls == i
        if inds.any():
            t = roi_align(feats[i], rois[inds, :], (oh, ow),
                          1 / featmap_strides[i], sample_num, backend)
            roi_feats[inds] = t
    return roi_feats


def arrff_bbox_feats(feats, rois, featmap_strides, out_size=7, sample_num=0,
                     finest_scale=56, backend="c"):
    """The (re-enabled) 3-region block of StandardRoIHead._bbox_forward,
    roi_heads/standard_roi_head.py:138-155: cat([ori, lw, lh], dim=1)."""
    n = len(featmap_strides)
    kw = dict(featmap_strides=featmap_strides, out_size=out_size,
              sample_num=sample_num, finest_scale=finest_scale,
              backend=backend)
    ori = single_roi_extractor(feats[:n], rois, **kw)
    lh_rois, lw_rois = get_adaptive_scale_rois(rois, 1)
    lh = single_roi_extractor(feats[:n], lh_rois, **kw)
    lw = single_roi_extractor(feats[:n], lw_rois, **kw)
    return torch.cat([ori, lw, lh], dim=1)


def region_boxes_and_levels(rois, num_levels, finest_scale=56):
    """[3,K,5] boxes in output-channel order (ori, lw, lh) and [3,K] levels."""
    lh_rois, lw_rois = get_adaptive_scale_rois(rois, 1)
    boxes = torch.stack([rois, lw_rois, lh_rois])
    lvls = torch.stack([map_roi_levels(b, num_levels, finest_scale)
                        for b in boxes])
    return boxes, lvls


def rff_gate(ori, a, b):
    """bbox_heads/multirois_bbox_head.py:175,182: ori + ori*(a+b), where a, b
    are already relu(conv(.)) (:172-173)."""
    return ori + ori * (a + b)


def rff_softmax_fuse(regions, logits):
    """bbox_heads/multirois_bbox_head.py:187-197 (commented in the shipped tree; the
    fusion drawn in the paper's figure): ws = softmax(logits, dim=1);
    out = r0 * ws[:, 0:1] + r1 * ws[:, 1:2] + r2 * ws[:, 2:3]."""
    ws = F.softmax(logits, dim=1)
    r0, r1, r2 = regions
    k, _, h, w = ws.shape
    return r0 * ws[:, 0, :, :].view(-1, 1, h, w) + r1 * ws[:, 1, :, :].view(-1, 1, h, w) + \
        r2 * ws[:, 2, :, :].view(-1, 1, h, w)


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule as the reference relies on it:
    Conv2d(bias=True) then ReLU unless act_cfg=None; parameters under .conv"""

    def __init__(self, cin, cout, k, padding=0, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, bias=True)
        self.act = act

    def forward(self, x):
        x = self.conv(x)
        return F.relu(x) if self.act else x


class MultiRoIsBBoxHead(nn.Module):
    """bbox_heads/multirois_bbox_head.py:13-251 with the MultiRoIsBBoxHead
    preset (:238-251): 2 shared FCs, no extra convs; BBoxHead ctor defaults
    bbox_head.py:19-66.  Loss/target code is out of scope."""

    def __init__(self, in_channels=256, fc_out_channels=1024, roi_feat_size=7,
                 num_classes=80, reg_class_agnostic=False):
        super().__init__()
        self.conv_out_channels = 256
        self.in_channels = in_channels
        self.hh_conv = ConvModule(in_channels, in_channels, 3, padding=1)
        self.wh_conv = ConvModule(in_channels, in_channels, 3, padding=1)
        self.final_conv = ConvModule(in_channels, in_channels, 3, padding=1)
        area = roi_feat_size * roi_feat_size
        self.shared_fcs = nn.ModuleList([
            nn.Linear(in_channels * area, fc_out_channels),
            nn.Linear(fc_out_channels, fc_out_channels)])
        self.fc_cls = nn.Linear(fc_out_channels, num_classes + 1)
        self.fc_reg = nn.Linear(
            fc_out_channels, 4 if reg_class_agnostic else 4 * num_classes)

    def forward(self, x):
        c = self.conv_out_channels
        ori = x[:, :c]                                  # :167
        lwh = F.relu(self.wh_conv(x[:, c:2 * c]))       # :168,:172
        lhh = F.relu(self.hh_conv(x[:, 2 * c:]))        # :169,:173
        ori_feats = ori * (lwh + lhh)                   # :175
        x_out = ori + ori_feats                         # :182
        x_out = F.relu(self.final_conv(x_out))          # :183
        x_out = x_out.flatten(1)                        # :203
        for fc in self.shared_fcs:
            x_out = F.relu(fc(x_out))                   # :204-205
        return self.fc_cls(x_out), self.fc_reg(x_out)   # :233-235


# --------------------------------------------------------------------------
# AR-FPN (WFPNDualSpatial)
# --------------------------------------------------------------------------
def wfpn_gather(inputs, refine_level=2):
    """necks/wfpn_dual_spatial.py:102-113."""
    size = inputs[refine_level].size()[2:]
    feats = []
    for i in range(len(inputs)):
        if i < refine_level:
            feats.append(F.adaptive_max_pool2d(inputs[i], output_size=size))
        else:
            feats.append(F.interpolate(inputs[i], size=size, mode="nearest"))
    return sum(feats) / len(feats)


def wfpn_apply(inputs, bsf, gate1, gate2):
    """necks/wfpn_dual_spatial.py:118-135.  gate1[i], gate2[i] are the raw
    Conv2d outputs of reduce_convs[i] / reduce_convs2[i] (bias included, before
    ConvModule's ReLU); the ReLU (mmcv default act) and tanh are applied here."""
    outs = []
    for i, x in enumerate(inputs):
        h, w = x.shape[2:]
        basic = torch.tanh(F.relu(gate1[i]))
        com = torch.tanh(F.relu(gate2[i]))
        att = F.interpolate(bsf, size=[h, w]) * (basic + com)
        outs.append(x + att)
    return tuple(outs)


def nonlocal_attention(theta, phi, g, use_scale=False, round_operands=None):
    """ops/non_local.py:78-101 between the 1x1 convolutions and conv_out: theta, phi, g are the
    [N, C, H, W] outputs of self.theta / self.phi / self.g; returns y [N, C, H, W] (:101).
    embedded_gaussian (:65-69) with the optional 1/sqrt(C) scale.  round_operands (a torch dtype)
    rounds the three operands first -- the arithmetic a reduced-precision tensor-core kernel is
    entitled to -- and computes everything else in fp32 like the reference."""
    n, c, h, w = theta.shape
    f = torch.float32
    if round_operands is not None:
        theta, phi, g = (t.to(round_operands) for t in (theta, phi, g))
    g_x = g.to(f).reshape(n, c, -1).permute(0, 2, 1)              # :83-84
    theta_x = theta.to(f).reshape(n, c, -1).permute(0, 2, 1)      # :87-89
    phi_x = phi.to(f).reshape(n, c, -1)                           # :92
    pw = torch.matmul(theta_x, phi_x)                             # :67
    if use_scale:
        pw = pw / theta_x.shape[-1] ** 0.5                        # :68-70
    pw = pw.softmax(dim=-1)                                       # :71
    y = torch.matmul(pw, g_x)                                     # :99
    return y.permute(0, 2, 1).contiguous().reshape(n, c, h, w)    # :101


class NonLocal2D(nn.Module):
    """ops/non_local.py:22-105 as configured by the neck
    (wfpn_dual_spatial.py:78-83): reduction=1, use_scale=False,
    embedded_gaussian."""

    def __init__(self, c):
        super().__init__()
        self.g = ConvModule(c, c, 1, act=False)
        self.theta = ConvModule(c, c, 1, act=False)
        self.phi = ConvModule(c, c, 1, act=False)
        self.conv_out = ConvModule(c, c, 1, act=False)
        for m in (self.g, self.theta, self.phi):        # :55-57
            nn.init.normal_(m.conv.weight, 0, 0.01)
            nn.init.constant_(m.conv.bias, 0)
        nn.init.constant_(self.conv_out.conv.weight, 0)  # :58-59
        nn.init.constant_(self.conv_out.conv.bias, 0)

    def forward(self, x):
        n, c, h, w = x.shape
        y = nonlocal_attention(self.theta(x), self.phi(x), self.g(x))
        return x + self.conv_out(y)


class WFPNDualSpatial(nn.Module):
    """necks/wfpn_dual_spatial.py:10-137."""

    def __init__(self, in_channels, num_levels, refine_level=2):
        super().__init__()
        self.num_levels = num_levels
        self.refine_level = refine_level
        self.reduce_convs = nn.ModuleList(
            [ConvModule(in_channels, 1, 3, padding=1) for _ in range(num_levels)])
        self.reduce_convs2 = nn.ModuleList(
            [ConvModule(in_channels, 1, 3, padding=1) for _ in range(num_levels)])
        self.refine = NonLocal2D(in_channels)

    def init_weights(self):                              # :94-97
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight, gain=1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        assert len(inputs) == self.num_levels
        ori_fe = wfpn_gather(inputs, self.refine_level)
        bsf = self.refine(ori_fe)
        g1 = [self.reduce_convs[i].conv(inputs[i]) for i in range(self.num_levels)]
        g2 = [self.reduce_convs2[i].conv(inputs[i]) for i in range(self.num_levels)]
        return wfpn_apply(inputs, bsf, g1, g2)


# --------------------------------------------------------------------------
# Synthetic inputs of SURVEY.md section 8(d)
# --------------------------------------------------------------------------
def synthetic_rois(K, img_w=1344, img_h=800, batch=1, seed=0,
                   smin=16.0, smax=600.0, order="interleaved"):
    """Centre uniform in the canvas, sqrt(area) log-uniform in [smin, smax],
    aspect log-uniform in [0.5, 2], clipped to the image, seeded.
    order: "interleaved" (RoI i -> image i % batch) or "image_major" (per-image
    blocks, what bbox2roi builds: core/bbox/transforms.py:51-59)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(K, 4, generator=g)
    cx = u[:, 0] * img_w
    cy = u[:, 1] * img_h
    s = torch.exp(np.log(smin) + u[:, 2] * (np.log(smax) - np.log(smin)))
    ar = torch.exp(np.log(0.5) + u[:, 3] * (np.log(2.0) - np.log(0.5)))
    w = s * torch.sqrt(ar)
    h = s / torch.sqrt(ar)
    x1 = (cx - w / 2).clamp(0, img_w - 1)
    y1 = (cy - h / 2).clamp(0, img_h - 1)
    x2 = (cx + w / 2).clamp(0, img_w - 1)
    y2 = (cy + h / 2).clamp(0, img_h - 1)
    x2 = torch.max(x2, x1 + 1.0)
    y2 = torch.max(y2, y1 + 1.0)
    if order == "image_major":
        b = ((torch.arange(K) * batch) // max(K, 1)).float()
    else:
        b = (torch.arange(K) % batch).float()
    return torch.stack([b, x1, y1, x2, y2], dim=1).float().contiguous()


def pyramid_shapes(img_h=800, img_w=1344, strides=(4, 8, 16, 32, 64)):
    """FPN level sizes: conv stride-2 chain of ceil halving (necks/fpn.py:202
    extra level = max_pool2d(k=1,s=2) -> ceil(n/2) as floor((n-1)/2)+1)."""
    shapes = []
    h, w = img_h, img_w
    s = 1
    for st in strides:
        while s < st:
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            s *= 2
        shapes.append((h, w))
    return shapes


def synthetic_pyramid(batch=1, channels=256, shapes=None, seed=0,
                      dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    shapes = shapes or pyramid_shapes()
    return [torch.randn(batch, channels, h, w, generator=g).to(dtype)
            for (h, w) in shapes]


# --------------------------------------------------------------------------
# The bench workload's training step, composed from the restated reference ops
# --------------------------------------------------------------------------
def reference_step(host, backend="c", channels=None, dy_full=None, refine_level=2, roi_levels=4):
    """One training step of the region-aware feature path on the inputs of
    arfe_b200.workload.host_inputs (plain dict of CPU tensors), in the
    reference's arithmetic -- what arfe_b200.workload.TrainStep.step computes:

      gathered = wfpn_gather(x)                       wfpn_dual_spatial.py:102-113
      y        = wfpn_apply(x, bsf, g1, g2)           :118-135
      F        = cat(ori, lw, lh) RoI features of y   standard_roi_head.py:138-155
      z        = ori + ori * (a + b)                  multirois_bbox_head.py:175,182
      backward from (z, lw, lh, gathered) with the incoming gradients
      (gz, glw, glh, gbsf); d x_l = d y_l + the gather's routed gradient.

    roi_levels: the extractor reads x[:num_inputs] with featmap_strides
    [4, 8, 16, 32] (configs/_base_/models/faster_rcnn_r50_fpn.py:44,
    standard_roi_head.py:140).
    channels: RoIAlign and the gate are independent per channel, so a checker
    may run the RoI part on a channel subset (all K RoIs, exact per element);
    the AR-FPN backward then needs d y over ALL channels (its gate gradients
    are channel sums): pass the tensor under test as dy_full, or leave it None
    to skip that part.  Returns a dict of tensors plus wall-clock seconds
    t_fpn / t_roi (the two CPU arms of bench.py)."""
    import time
    strides = list(host.get("strides", (4, 8, 16, 32, 64)))
    P = host.get("out_size", 7)
    f32 = lambda t: t.detach().float().contiguous()
    C = host["a"].shape[1]
    S = list(range(C)) if channels is None else list(channels)
    c = len(S)
    out = {}
    t0 = time.perf_counter()
    x = [f32(t).requires_grad_(True) for t in host["x"]]
    bsf = f32(host["bsf"]).requires_grad_(True)
    g1 = [f32(t).requires_grad_(True) for t in host["g1"]]
    g2 = [f32(t).requires_grad_(True) for t in host["g2"]]
    gathered = wfpn_gather(x, refine_level)
    y = wfpn_apply(x, bsf, g1, g2)
    t1 = time.perf_counter()
    yd = [t.detach()[:, S].contiguous().requires_grad_(True) for t in y]
    F_ = arrff_bbox_feats(yd, f32(host["rois"]), strides[:roi_levels], out_size=P, backend=backend)
    a = f32(host["a"])[:, S].contiguous().requires_grad_(True)
    b = f32(host["b"])[:, S].contiguous().requires_grad_(True)
    ori = F_[:, :c]
    ori.retain_grad()
    z = rff_gate(ori, a, b)
    sub = lambda k: f32(host[k])[:, S].contiguous()
    torch.autograd.backward([z, F_[:, c:2 * c], F_[:, 2 * c:]], [sub("gz"), sub("glw"), sub("glh")])
    dy = [t.grad if t.grad is not None else torch.zeros_like(t) for t in yd]
    t2 = time.perf_counter()
    out.update(gathered=gathered.detach(), y=[t.detach() for t in y], F=F_.detach(), z=z.detach(),
               d_ori=ori.grad.detach(), d_ab=a.grad.detach(), dy=dy, channels=S)
    t3 = t2
    if channels is None or dy_full is not None:
        dyf = dy if dy_full is None else [f32(t) for t in dy_full]
        torch.autograd.backward(list(y) + [gathered], list(dyf) + [f32(host["gbsf"])])
        t3 = time.perf_counter()
        out.update(dbsf=bsf.grad, dg1=[t.grad for t in g1], dg2=[t.grad for t in g2],
                   dx=[t.grad for t in x])
    out["t_fpn"] = (t1 - t0) + (t3 - t2)
    out["t_roi"] = t2 - t1
    return out
