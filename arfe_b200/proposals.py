"""Proposal side of the path with the reference's call surface: ``nms`` and
``batched_nms`` (mmdet/ops/nms/nms_wrapper.py:7-66, :119-157) and ``bbox2roi``
(mmdet/core/bbox/transforms.py:41-60), on arfe_nms / arfe_bbox2roi.

The reference builds the suppression bitmask on the device, copies it to the
host and sweeps it there; here the sweep stays on the device, so the only
device->host traffic of ``nms`` is the 4-byte survivor count that sizes the
result (the result tensors are data dependent in the reference's API).
"""
import ctypes

import torch

from . import _lib as L


def nms(dets, iou_thr, device_id=None):
    """dets: [N, 5] CUDA tensor (x1, y1, x2, y2, score).  Returns (dets[inds], inds) with inds in
    descending-score order, like mmdet.ops.nms."""
    if not torch.is_tensor(dets):
        raise TypeError("arfe_b200.nms takes a CUDA tensor (the reference's numpy path is CPU code)")
    L.require_cuda(dets)
    if dets.dim() != 2 or dets.size(1) != 5:
        raise AssertionError("dets must be [N, 5] = (x1, y1, x2, y2, score)")
    n = dets.size(0)
    if n == 0:
        return dets, dets.new_zeros(0, dtype=torch.long)
    d32 = dets.detach().float()
    order = d32[:, 4].sort(0, descending=True)[1]          # nms_kernel.cu:79-80
    ds = d32.index_select(0, order).contiguous()
    lib = L.lib()
    nbytes = lib.arfe_nms_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dets.device)
    keep = torch.empty(n, dtype=torch.long, device=dets.device)
    cnt = torch.empty(1, dtype=torch.int32, device=dets.device)
    L.check(lib.arfe_nms(ds.data_ptr(), n, float(iou_thr), ws.data_ptr(), nbytes, keep.data_ptr(),
                         cnt.data_ptr(), L.stream_ptr(dets.device)), "arfe_nms")
    inds = order[keep[:int(cnt.item())]]
    return dets[inds, :], inds


def batched_nms(bboxes, scores, inds, nms_cfg, class_agnostic=False):
    """mmdet/ops/nms/nms_wrapper.py:119-157: NMS per cluster (class / level) through a
    per-cluster coordinate offset."""
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        boxes_for_nms = bboxes
    else:
        max_coordinate = bboxes.max()
        offsets = inds.to(bboxes) * (max_coordinate + 1)
        boxes_for_nms = bboxes + offsets[:, None]
    nms_type = cfg.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError(f"nms type {nms_type!r}: only hard NMS is provided (soft_nms is CPU code in the reference)")
    dets, keep = nms(torch.cat([boxes_for_nms, scores[:, None]], -1), **cfg)
    return torch.cat([bboxes[keep], dets[:, -1:]], -1), keep


def bbox2roi(bbox_list):
    """list of [n_i, >=4] CUDA tensors -> [sum n_i, 5] (img_id, x1, y1, x2, y2), image-major."""
    if len(bbox_list) == 0:
        return torch.zeros((0, 5))
    L.require_cuda(*bbox_list)
    dev = bbox_list[0].device
    boxes = [b.detach().float().contiguous() for b in bbox_list]
    counts = [int(b.size(0)) for b in boxes]
    cols = [int(b.size(1)) if b.dim() == 2 and b.size(0) > 0 else 4 for b in boxes]
    rois = torch.empty((sum(counts), 5), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * len(boxes))(*[b.data_ptr() if c > 0 else None for b, c in zip(boxes, counts)])
    L.check(L.lib().arfe_bbox2roi(ptrs, L.int_array(counts), L.int_array(cols), len(boxes), rois.data_ptr(),
                                  L.stream_ptr(dev)), "arfe_bbox2roi")
    return rois.to(bbox_list[0].dtype)
