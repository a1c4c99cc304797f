"""torch.autograd Functions over the C ABI (include/arfe_b200.h).

PyTorch is plumbing here: it owns device memory and the stream, the math runs
in libarfe_b200.so.  Every op raises on non-CUDA tensors.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _lib as L


def _prep_feats(feats):
    """Bring all levels to one dtype/layout (the first level's)."""
    L.require_cuda(*feats)
    layout = L.layout_of(feats[0])
    return [L.as_layout(f, layout) for f in feats], layout, L.dtype_code(feats[0])


def _zeros(shape, dtype, device, layout):
    mf = torch.channels_last if layout == L.ARFE_NHWC else torch.contiguous_format
    return torch.empty(shape, dtype=dtype, device=device, memory_format=mf).zero_()


def _prep_rois(rois):
    L.require_cuda(rois)
    if rois.dim() != 2 or rois.size(1) != 5:
        raise AssertionError("rois must be [K, 5] = (batch_idx, x1, y1, x2, y2)")
    return rois.detach().float().contiguous()


def _vec_ok(C, dtype):
    return C % (4 if dtype == torch.float32 else 8) == 0


_GEO = {}


def _geo(Hs, Ws, scales):
    """ctypes copies of the level sizes / scales, built once per pyramid geometry."""
    key = (tuple(Hs), tuple(Ws), tuple(float(s) for s in scales))
    g = _GEO.get(key)
    if g is None:
        if len(_GEO) > 64:
            _GEO.clear()
        g = _GEO[key] = (L.int_array(Hs), L.int_array(Ws), L.float_array(scales))
    return g


_SIDE = {}


def _side_stream(device):
    s = _SIDE.get(device.index)
    if s is None:
        s = _SIDE[device.index] = torch.cuda.Stream(device)
    return s


class _Plan:
    """RoI plan of one forward (region tables in a device workspace) and, when a
    backward will follow, the pull kernel's tile bins -- built on a side stream
    while the forward kernels run on the caller's stream; the backward joins on
    `ev_bin` (plan_ready = 2)."""

    def __init__(self, geo_args, K, regions, nlev, B, Hs_c, Ws_c, device, need_bins, split):
        lib = L.lib()
        self.nbytes = lib.arfe_roi_plan_bytes(K, regions, nlev, B, Hs_c, Ws_c)
        self.ws = torch.empty(self.nbytes + 256, dtype=torch.uint8, device=device)
        self.ptr = (self.ws.data_ptr() + 255) // 256 * 256
        self.ev_bin = None
        main = torch.cuda.current_stream(device)
        L.check(lib.arfe_roi_plan_build(*geo_args, self.ptr, self.nbytes, c_stream(main)),
                "arfe_roi_plan_build")
        self._bins = (geo_args, int(split)) if need_bins else None

    def fork_bins(self, device):
        """Called right after the forward's launch: the tile bins are built on the side stream
        next to whatever follows the extraction (the head's convolutions), not next to the ring
        forward, which a second latency-bound kernel slows down."""
        if self._bins is None or torch.cuda.is_current_stream_capturing():
            return
        geo_args, split = self._bins
        main, side = torch.cuda.current_stream(device), _side_stream(device)
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        self.ws.record_stream(side)
        L.check(L.lib().arfe_roi_pull_bin(*geo_args, split, self.ptr, self.nbytes, c_stream(side)),
                "arfe_roi_pull_bin")
        self.ev_bin = torch.cuda.Event()
        self.ev_bin.record(side)

    def ready_for_backward(self, device):
        """plan_ready code for the backward: 2 = plan and bins are there (after joining the
        side stream), 1 = plan only."""
        if self.ev_bin is None:
            return 1
        torch.cuda.current_stream(device).wait_event(self.ev_bin)
        return 2


def c_stream(stream):
    import ctypes
    return ctypes.c_void_p(stream.cuda_stream)


class _RoIFuseFunction(Function):
    """Fused region generation + level map + multi-level RoIAlign (+ cat).

    regions=1 reproduces SingleRoIExtractor.forward
    (roi_extractors/single_level.py:109-152); regions=3 reproduces the AR-RFF
    block of StandardRoIHead._bbox_forward (standard_roi_head.py:138-155).
    Channels-last features take the fast path: 128-bit gathers forward, the
    atomic-free deterministic "pull" kernel backward.
    """

    @staticmethod
    def forward(ctx, rois, out_size, spatial_scales, sample_num, regions, facs,
                finest_scale, out_channels_last, *feats):
        oh, ow = _pair(out_size)
        feats, layout, dt = _prep_feats(feats)
        rois = _prep_rois(rois)
        B, C = feats[0].shape[:2]
        Hs = [f.shape[2] for f in feats]
        Ws = [f.shape[3] for f in feats]
        K = rois.size(0)
        fast = layout == L.ARFE_NHWC and _vec_ok(C, feats[0].dtype)
        out_cl = bool(out_channels_last) and fast
        out = torch.empty((K, regions * C, oh, ow), dtype=feats[0].dtype, device=feats[0].device,
                          memory_format=torch.channels_last if out_cl else torch.contiguous_format)
        ctx.meta = (oh, ow, tuple(spatial_scales), sample_num, regions, facs,
                    finest_scale, layout, dt, B, C, Hs, Ws, feats[0].dtype, fast)
        ctx.save_for_backward(rois)
        ctx.plan = None
        Hs_c, Ws_c, sc_c = _geo(Hs, Ws, spatial_scales)
        if K > 0 and out_cl:
            # channels-last in and out: plan + ring kernel; the plan (region tables in
            # a device workspace) is kept for the backward, whose tile bins are built on
            # a side stream while the forward runs
            need_bwd = any(f.requires_grad for f in feats)
            geo_args = (Hs_c, Ws_c, sc_c, len(feats), B, C, rois.data_ptr(), K, regions, float(facs),
                        oh, ow, int(sample_num), float(finest_scale), dt)
            plan = _Plan(geo_args, K, regions, len(feats), B, Hs_c, Ws_c, out.device, need_bwd, False)
            rc = L.lib().arfe_roi_fuse_forward_plan(
                L.ptr_array(feats), Hs_c, Ws_c, sc_c, len(feats), B, C, rois.data_ptr(),
                K, regions, float(facs), oh, ow, int(sample_num), float(finest_scale), dt,
                out.data_ptr(), None, None, plan.ptr, plan.nbytes, 1, L.stream_ptr(out.device))
            L.check(rc, "arfe_roi_fuse_forward_plan")
            if need_bwd:
                plan.fork_bins(out.device)
                ctx.plan = plan
        elif K > 0:
            rc = L.lib().arfe_roi_fuse_forward(
                L.ptr_array(feats), Hs_c, Ws_c, sc_c, len(feats), B, C, rois.data_ptr(),
                K, regions, float(facs), oh, ow, int(sample_num),
                float(finest_scale), dt, layout, L.ARFE_NHWC if out_cl else L.ARFE_NCHW,
                out.data_ptr(), None, None, L.stream_ptr(out.device))
            L.check(rc, "arfe_roi_fuse_forward")
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (rois,) = ctx.saved_tensors
        (oh, ow, scales, sample_num, regions, facs, finest_scale, layout, dt, B,
         C, Hs, Ws, fdtype, fast) = ctx.meta
        nlev = len(Hs)
        NP = 8  # non-tensor inputs of forward()
        if not any(ctx.needs_input_grad[NP:]):
            return (None,) * (NP + nlev)
        K = rois.size(0)
        dev = grad_out.device
        g = grad_out if grad_out.dtype == fdtype else grad_out.to(fdtype)
        lib = L.lib()
        if fast and K > 0:
            # pull kernel: channels-last dout, every gradient element written once
            g = g.contiguous(memory_format=torch.channels_last)
            Hs_c, Ws_c, sc_c = _geo(Hs, Ws, scales)
            if ctx.plan is not None:
                ws, ws_ptr, nbytes = ctx.plan.ws, ctx.plan.ptr, ctx.plan.nbytes
                ready = ctx.plan.ready_for_backward(dev)
            else:
                nbytes = lib.arfe_roi_plan_bytes(K, regions, nlev, B, Hs_c, Ws_c)
                ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
                ws_ptr = (ws.data_ptr() + 255) // 256 * 256
                ready = 0
            dfeats = [torch.empty((B, C, Hs[l], Ws[l]), dtype=torch.float32, device=dev,
                                  memory_format=torch.channels_last) for l in range(nlev)]
            rc = lib.arfe_roi_fuse_backward_pull(
                g.data_ptr(), Hs_c, Ws_c, sc_c, nlev, B,
                C, rois.data_ptr(), K, regions, float(facs), oh, ow, int(sample_num),
                float(finest_scale), dt, L.ptr_array(dfeats), ws_ptr, nbytes, ready,
                L.stream_ptr(dev))
            L.check(rc, "arfe_roi_fuse_backward_pull")
        else:
            dfeats = [_zeros((B, C, Hs[l], Ws[l]), torch.float32, dev, layout)
                      for l in range(nlev)]
            if K > 0:
                g_cl = L.layout_of(g) == L.ARFE_NHWC
                g = g.contiguous(memory_format=torch.channels_last) if g_cl else g.contiguous()
                rc = lib.arfe_roi_fuse_backward(
                    g.data_ptr(), L.ARFE_NHWC if g_cl else L.ARFE_NCHW, L.int_array(Hs),
                    L.int_array(Ws), L.float_array(scales), nlev, B, C, rois.data_ptr(), K,
                    regions, float(facs), oh, ow, int(sample_num), float(finest_scale), dt,
                    layout, L.ptr_array(dfeats), L.stream_ptr(dev))
                L.check(rc, "arfe_roi_fuse_backward")
        grads = tuple(d if fdtype == torch.float32 else d.to(fdtype) for d in dfeats)
        return (None,) * NP + grads


def roi_fuse(feats, rois, out_size, spatial_scales, sample_num=0, regions=3,
             facs=1.0, finest_scale=56, out_channels_last=False):
    """[K, regions*C, oh, ow]; channel blocks (ori, lw, lh) when regions=3.
    out_channels_last=True (channels-last features only) returns the tensor in
    torch.channels_last memory format -- what cuDNN's 3x3 convs prefer."""
    return _RoIFuseFunction.apply(rois, out_size, tuple(spatial_scales),
                                  sample_num, regions, facs, finest_scale,
                                  out_channels_last, *feats)


class _RoIFuseSplitFunction(Function):
    """Same extraction as _RoIFuseFunction for channels-last features, with the
    regions as SEPARATE channels-last tensors (ori, lw, lh) instead of
    torch.cat([ori, lw, lh], 1) (standard_roi_head.py:155): the head convolves
    each region on its own (multirois_bbox_head.py:167-173), so the cat and the
    slices (and, backward, the zero-fill + slice-assign autograd inserts for
    them) are pure copies.  Forward: plan + ring kernel; backward: pull kernel
    reading the three incoming gradients in place, reusing the plan."""

    @staticmethod
    def forward(ctx, rois, out_size, spatial_scales, sample_num, regions, facs,
                finest_scale, *feats):
        oh, ow = _pair(out_size)
        feats, layout, dt = _prep_feats(feats)
        rois = _prep_rois(rois)
        B, C = feats[0].shape[:2]
        if layout != L.ARFE_NHWC or not _vec_ok(C, feats[0].dtype):
            raise RuntimeError("roi_fuse_split needs channels-last features with C % 4 (fp32) / 8 (bf16) == 0")
        Hs = [f.shape[2] for f in feats]
        Ws = [f.shape[3] for f in feats]
        K = rois.size(0)
        dev = feats[0].device
        outs = [torch.empty((K, C, oh, ow), dtype=feats[0].dtype, device=dev,
                            memory_format=torch.channels_last) for _ in range(regions)]
        ctx.meta = (oh, ow, tuple(spatial_scales), sample_num, regions, facs, finest_scale, dt,
                    B, C, Hs, Ws, feats[0].dtype)
        ctx.save_for_backward(rois)
        ctx.plan = None
        if K > 0:
            lib = L.lib()
            Hs_c, Ws_c, sc_c = _geo(Hs, Ws, spatial_scales)
            need_bwd = any(f.requires_grad for f in feats)
            geo_args = (Hs_c, Ws_c, sc_c, len(feats), B, C, rois.data_ptr(), K, regions, float(facs),
                        oh, ow, int(sample_num), float(finest_scale), dt)
            plan = _Plan(geo_args, K, regions, len(feats), B, Hs_c, Ws_c, dev, need_bwd, True)
            rc = lib.arfe_roi_fuse_forward_plan_split(
                L.ptr_array(feats), Hs_c, Ws_c, sc_c, len(feats), B, C, rois.data_ptr(), K, regions,
                float(facs), oh, ow, int(sample_num), float(finest_scale), dt, L.ptr_array(outs),
                plan.ptr, plan.nbytes, 1, L.stream_ptr(dev))
            L.check(rc, "arfe_roi_fuse_forward_plan_split")
            if need_bwd:
                plan.fork_bins(dev)
                ctx.plan = plan
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grad_outs):
        (rois,) = ctx.saved_tensors
        (oh, ow, scales, sample_num, regions, facs, finest_scale, dt, B, C, Hs, Ws, fdtype) = ctx.meta
        nlev = len(Hs)
        NP = 7  # non-tensor inputs of forward()
        if not any(ctx.needs_input_grad[NP:]):
            return (None,) * (NP + nlev)
        K = rois.size(0)
        dev = rois.device
        lib = L.lib()
        gs = []
        for g in grad_outs:
            if g is None:
                g = torch.zeros((K, C, oh, ow), dtype=fdtype, device=dev)
            g = g if g.dtype == fdtype else g.to(fdtype)
            gs.append(g.contiguous(memory_format=torch.channels_last))
        dfeats = [torch.empty((B, C, Hs[l], Ws[l]), dtype=torch.float32, device=dev,
                              memory_format=torch.channels_last) for l in range(nlev)]
        if K > 0:
            Hs_c, Ws_c, sc_c = _geo(Hs, Ws, scales)
            if ctx.plan is not None:
                ws, ws_ptr, nbytes = ctx.plan.ws, ctx.plan.ptr, ctx.plan.nbytes
                ready = ctx.plan.ready_for_backward(dev)
            else:
                nbytes = lib.arfe_roi_plan_bytes(K, regions, nlev, B, Hs_c, Ws_c)
                ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
                ws_ptr = (ws.data_ptr() + 255) // 256 * 256
                ready = 0
            rc = lib.arfe_roi_fuse_backward_pull_split(
                L.ptr_array(gs), Hs_c, Ws_c, sc_c, nlev, B, C,
                rois.data_ptr(), K, regions, float(facs), oh, ow, int(sample_num), float(finest_scale),
                dt, L.ptr_array(dfeats), ws_ptr, nbytes, ready, L.stream_ptr(dev))
            L.check(rc, "arfe_roi_fuse_backward_pull_split")
        else:
            for d in dfeats:
                d.zero_()
        grads = tuple(d if fdtype == torch.float32 else d.to(fdtype) for d in dfeats)
        return (None,) * NP + grads


def roi_fuse_split(feats, rois, out_size, spatial_scales, sample_num=0, regions=3,
                   facs=1.0, finest_scale=56):
    """The `regions` region tensors (ori, lw, lh), each [K, C, oh, ow] in
    torch.channels_last -- torch.cat(roi_fuse_split(...), 1) == roi_fuse(...).
    Channels-last features only."""
    return _RoIFuseSplitFunction.apply(rois, out_size, tuple(spatial_scales), sample_num,
                                       regions, facs, finest_scale, *feats)


def roi_fuse_debug(rois, Hs, Ws, spatial_scales, out_size=7, sample_num=0,
                   regions=3, facs=1.0, finest_scale=56, max_grid=16):
    """Boxes, levels, grids and bilinear taps exactly as the kernels compute
    them (arfe_roi_fuse_taps) -- parity instrumentation."""
    oh, ow = _pair(out_size)
    rois = _prep_rois(rois)
    dev = rois.device
    K, R = rois.size(0), regions
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    o = dict(
        lvl=torch.zeros((R, K), **i32), grid=torch.zeros((R, K, 2), **i32),
        boxes=torch.zeros((R, K, 5), **f32),
        ylo=torch.zeros((R, K, oh, max_grid), **i32), yhi=torch.zeros((R, K, oh, max_grid), **i32),
        ywl=torch.zeros((R, K, oh, max_grid), **f32), ywh=torch.zeros((R, K, oh, max_grid), **f32),
        xlo=torch.zeros((R, K, ow, max_grid), **i32), xhi=torch.zeros((R, K, ow, max_grid), **i32),
        xwl=torch.zeros((R, K, ow, max_grid), **f32), xwh=torch.zeros((R, K, ow, max_grid), **f32))
    if K > 0:
        rc = L.lib().arfe_roi_fuse_taps(
            L.int_array(Hs), L.int_array(Ws), L.float_array(spatial_scales),
            len(Hs), rois.data_ptr(), K, R, float(facs), oh, ow, int(sample_num),
            float(finest_scale), max_grid, o["lvl"].data_ptr(), o["grid"].data_ptr(),
            o["boxes"].data_ptr(), o["ylo"].data_ptr(), o["yhi"].data_ptr(),
            o["ywl"].data_ptr(), o["ywh"].data_ptr(), o["xlo"].data_ptr(),
            o["xhi"].data_ptr(), o["xwl"].data_ptr(), o["xwh"].data_ptr(),
            L.stream_ptr(dev))
        L.check(rc, "arfe_roi_fuse_taps")
    return o


class RoIAlignFunction(Function):
    """mmdet/ops/roi_align/roi_align.py:9-73 over arfe_roi_align_forward /
    arfe_roi_align_backward (the twins of roi_align_ext.forward_v2/backward_v2)."""

    @staticmethod
    def forward(ctx, features, rois, out_size, spatial_scale, sample_num=0,
                aligned=True):
        out_h, out_w = _pair(out_size)
        assert isinstance(out_h, int) and isinstance(out_w, int)
        if not aligned:
            raise NotImplementedError(
                "aligned=False is the legacy v1 RoIAlign, which does not build "
                "on current toolchains (SURVEY.md section 2); only aligned=True "
                "is provided")
        (features,), layout, dt = _prep_feats([features])
        rois = _prep_rois(rois)
        B, C, H, W = features.shape
        K = rois.size(0)
        out = features.new_empty((K, C, out_h, out_w))
        ctx.meta = (out_h, out_w, float(spatial_scale), int(sample_num), layout,
                    dt, (B, C, H, W), features.dtype)
        ctx.save_for_backward(rois)
        if K > 0:
            rc = L.lib().arfe_roi_align_forward(
                features.data_ptr(), rois.data_ptr(), float(spatial_scale), out_h,
                out_w, int(sample_num), 1, B, C, H, W, K, dt, layout,
                out.data_ptr(), L.stream_ptr(out.device))
            L.check(rc, "arfe_roi_align_forward")
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        out_h, out_w, scale, sample_num, layout, dt, (B, C, H, W), fdtype = ctx.meta
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        grad_input = _zeros((B, C, H, W), torch.float32, grad_output.device, layout)
        K = rois.size(0)
        if K > 0:
            g = grad_output.contiguous()
            if g.dtype != fdtype:
                g = g.to(fdtype)
            rc = L.lib().arfe_roi_align_backward(
                g.data_ptr(), rois.data_ptr(), scale, out_h, out_w, B, C, H, W, K,
                sample_num, 1, dt, layout, grad_input.data_ptr(),
                L.stream_ptr(g.device))
            L.check(rc, "arfe_roi_align_backward")
        if fdtype != torch.float32:
            grad_input = grad_input.to(fdtype)
        return grad_input, None, None, None, None, None


roi_align = RoIAlignFunction.apply


class _Split3(Function):
    """x[:, :C], x[:, C:2C], x[:, 2C:] (multirois_bbox_head.py:167-169) whose
    backward is ONE concatenation instead of three zero-filled full-size
    gradients plus two additions."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.c = c
        ctx.shape = x.shape
        return x[:, :c], x[:, c:2 * c], x[:, 2 * c:]

    @staticmethod
    def backward(ctx, g0, g1, g2):
        c, shape = ctx.c, ctx.shape
        ref = next(g for g in (g0, g1, g2) if g is not None)
        parts = []
        for g, n in ((g0, c), (g1, c), (g2, shape[1] - 2 * c)):
            parts.append(g if g is not None else
                         ref.new_zeros((shape[0], n) + tuple(shape[2:])))
        return torch.cat(parts, dim=1), None


def split3(x, c):
    return _Split3.apply(x, c)


def _gate_rows(t):
    """How the gate kernel can walk a [K, C, H, W] tensor in place:
    ("nchw", rows=K, n=C*H*W, row stride) when every RoI's block is NCHW-dense
    (a channel slice of the NCHW cat tensor qualifies), ("nhwc", rows=K*H*W,
    n=C, bin stride) when it is channels-last with a uniform bin stride (a
    region tensor of roi_fuse_split, or a channel slice of the channels-last
    cat tensor), else None.  Holds for K == 1 too (stride(0) is then free)."""
    if t.dim() != 4:
        return None
    K, C, H, W = t.shape
    s0, s1, s2, s3 = t.stride()
    inner_nchw = (C == 1 or s1 == H * W) and (H == 1 or s2 == W) and (W == 1 or s3 == 1)
    if inner_nchw and (K == 1 or s0 >= C * H * W):
        return "nchw", K, C * H * W, (s0 if K > 1 else C * H * W)
    cs = s3 if W > 1 else (s2 if H > 1 else (s0 if K > 1 else C))  # elements between consecutive bins
    inner_nhwc = (C == 1 or s1 == 1) and (W == 1 or s3 == cs) and (H == 1 or s2 == W * cs) and cs >= C
    if inner_nhwc and (K == 1 or s0 == H * W * cs):
        return "nhwc", K * H * W, C, cs
    return None


class _RFFGateFunction(Function):
    """out = ori + ori*(a+b), multirois_bbox_head.py:175,182.  ori is read in
    place from wherever it lives (a slice of the cat tensor or a region tensor
    of its own); a, b, out and the gradients follow its memory layout."""

    @staticmethod
    def forward(ctx, ori, a, b):
        L.require_cuda(ori, a, b)
        dt = L.dtype_code(ori)
        view = _gate_rows(ori)
        if view is None:
            ori = ori.contiguous()
            view = _gate_rows(ori) or ("nchw", ori.size(0), ori[0].numel() if ori.size(0) else 1,
                                       ori[0].numel() if ori.size(0) else 1)
        kind, rows, n, stride = view
        mf = torch.channels_last if kind == "nhwc" else torch.contiguous_format
        dense = (lambda t: t.to(ori.dtype).contiguous(memory_format=mf)) if ori.dim() == 4 else \
            (lambda t: t.to(ori.dtype).contiguous())
        a_in, b_in = a, b
        a, b = dense(a), dense(b)
        out = torch.empty(ori.shape, dtype=ori.dtype, device=ori.device,
                          **({"memory_format": mf} if ori.dim() == 4 else {}))
        if rows > 0:
            rc = L.lib().arfe_rff_gate_forward(
                ori.data_ptr(), stride, a.data_ptr(), b.data_ptr(),
                out.data_ptr(), rows, n, dt, L.stream_ptr(ori.device))
            L.check(rc, "arfe_rff_gate_forward")
        ctx.save_for_backward(ori, a, b)
        ctx.meta = (rows, n, stride, dt, mf, a_in.dtype, b_in.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        ori, a, b = ctx.saved_tensors
        rows, n, stride, dt, mf, a_dtype, b_dtype = ctx.meta
        kw = {"memory_format": mf} if ori.dim() == 4 else {}
        g = g.to(ori.dtype).contiguous(**kw)
        d_ori = torch.empty(ori.shape, dtype=ori.dtype, device=ori.device, **kw)
        d_ab = torch.empty(ori.shape, dtype=ori.dtype, device=ori.device, **kw)
        if rows > 0:
            rc = L.lib().arfe_rff_gate_backward(
                g.data_ptr(), ori.data_ptr(), stride, a.data_ptr(), b.data_ptr(),
                d_ori.data_ptr(), n, d_ab.data_ptr(), rows, n, dt,
                L.stream_ptr(g.device))
            L.check(rc, "arfe_rff_gate_backward")
        return d_ori, d_ab.to(a_dtype), d_ab.to(b_dtype)


def rff_gate(ori, a, b):
    return _RFFGateFunction.apply(ori, a, b)


def _kbc_strides(t):
    """(k, bin, channel) element strides of a [K, C, H, W] tensor whose bins are evenly spaced
    (NCHW-dense or channels-last, possibly a channel slice), else None."""
    K, C, H, W = t.shape
    s0, s1, s2, s3 = t.stride()
    bs = s3 if W > 1 else (s2 if H > 1 else 1)
    if H > 1 and W > 1 and s2 != W * s3:
        return None
    return (s0 if K > 1 else C * H * W * max(bs, 1), bs, s1 if C > 1 else 1)


class _RFFSoftmaxFuseFunction(Function):
    """Softmax-over-regions fusion (multirois_bbox_head.py:187-197, the commented variant
    of the paper's figure): out = sum_j r_j * softmax(logits, 1)[:, j]."""

    @staticmethod
    def forward(ctx, logits, r0, r1, r2):
        L.require_cuda(logits, r0, r1, r2)
        K, C, H, W = r0.shape
        dt = L.dtype_code(r0)
        regs = [r0, r1, r2]
        st = _kbc_strides(r0)
        if st is None or any(r.dtype != r0.dtype or _kbc_strides(r) != st for r in regs):
            mf = torch.channels_last if L.layout_of(r0) == L.ARFE_NHWC else torch.contiguous_format
            regs = [r.to(r0.dtype).contiguous(memory_format=mf) for r in regs]
            st = _kbc_strides(regs[0])
        mf = torch.channels_last if st[2] == 1 and C > 1 else torch.contiguous_format
        logits_c = logits.to(r0.dtype).contiguous()            # [K, 3, H, W]: strides (3 PP, PP, 1)
        out = torch.empty((K, C, H, W), dtype=r0.dtype, device=r0.device, memory_format=mf)
        PP = H * W
        lst = (3 * PP, PP, 1)
        if K > 0:
            rc = L.lib().arfe_rff_softmax_fuse_forward(
                L.ptr_array(regs), L.i64_array(st), logits_c.data_ptr(), L.i64_array(lst), out.data_ptr(),
                L.i64_array(_kbc_strides(out)), K, PP, C, dt, L.stream_ptr(out.device))
            L.check(rc, "arfe_rff_softmax_fuse_forward")
        ctx.save_for_backward(logits_c, *regs)
        ctx.meta = (st, lst, mf, dt, logits.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        logits_c, r0, r1, r2 = ctx.saved_tensors
        st, lst, mf, dt, ldtype = ctx.meta
        K, C, H, W = r0.shape
        g = g.to(r0.dtype).contiguous(memory_format=mf)
        dregs = [torch.empty((K, C, H, W), dtype=r0.dtype, device=g.device, memory_format=mf) for _ in range(3)]
        dlog = torch.empty_like(logits_c)
        if K > 0:
            rc = L.lib().arfe_rff_softmax_fuse_backward(
                g.data_ptr(), L.ptr_array([r0, r1, r2]), L.i64_array(st), logits_c.data_ptr(), L.i64_array(lst),
                L.i64_array(_kbc_strides(g)), L.ptr_array(dregs), dlog.data_ptr(), K, H * W, C, dt,
                L.stream_ptr(g.device))
            L.check(rc, "arfe_rff_softmax_fuse_backward")
        return (dlog.to(ldtype),) + tuple(dregs)


def rff_softmax_fuse(regions, logits):
    """regions: the three region tensors (r0, r1, r2), each [K, C, h, w] (e.g. from roi_fuse_split), or
    the concatenated [K, 3C, h, w] tensor; logits: [K, 3, h, w].  Returns [K, C, h, w]."""
    if torch.is_tensor(regions):
        c = regions.shape[1] // 3
        regions = (regions[:, :c], regions[:, c:2 * c], regions[:, 2 * c:])
    return _RFFSoftmaxFuseFunction.apply(logits, *regions)


class _FPNGateConvFunction(Function):
    """g1_l, g2_l = the two C -> 1 3x3 convolutions of every level (wfpn_dual_spatial.py:120-121),
    x read once for both (arfe_fpn_gate_conv_forward); the backward of all of them -- d x, d weights,
    d biases -- is again one pass over x (arfe_fpn_gate_conv_backward)."""

    @staticmethod
    def forward(ctx, nlev, *tensors):
        feats = tensors[:nlev]
        w1, b1 = tensors[nlev:2 * nlev], tensors[2 * nlev:3 * nlev]
        w2, b2 = tensors[3 * nlev:4 * nlev], tensors[4 * nlev:5 * nlev]
        feats, layout, dt = _prep_feats(feats)
        if layout != L.ARFE_NHWC:
            raise RuntimeError("fpn_gate_conv needs channels-last feature maps")
        B, C = feats[0].shape[:2]
        Hs = [f.shape[2] for f in feats]
        Ws = [f.shape[3] for f in feats]
        dev = feats[0].device
        f32 = lambda ts: [t.detach().float().contiguous() for t in ts]
        w1f, b1f, w2f, b2f = f32(w1), f32(b1), f32(w2), f32(b2)
        g1 = [torch.empty((B, 1, h, w), dtype=feats[0].dtype, device=dev) for h, w in zip(Hs, Ws)]
        g2 = [torch.empty_like(t) for t in g1]
        if B > 0:
            lib = L.lib()
            nbytes = lib.arfe_fpn_gate_conv_workspace_bytes(nlev, B, L.int_array(Hs), L.int_array(Ws))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rc = lib.arfe_fpn_gate_conv_forward(
                L.ptr_array(feats), L.ptr_array(w1f), L.ptr_array(b1f), L.ptr_array(w2f), L.ptr_array(b2f),
                L.int_array(Hs), L.int_array(Ws), nlev, B, C, dt, layout, ws.data_ptr(), nbytes,
                L.ptr_array(g1), L.ptr_array(g2), L.stream_ptr(dev))
            L.check(rc, "arfe_fpn_gate_conv_forward")
        ctx.nlev = nlev
        ctx.save_for_backward(*feats, *w1, *w2)
        return tuple(g1) + tuple(g2)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        n = ctx.nlev
        saved = ctx.saved_tensors
        feats, w1, w2 = list(saved[:n]), saved[n:2 * n], saved[2 * n:3 * n]
        need = ctx.needs_input_grad
        B, C = feats[0].shape[:2]
        Hs = [f.shape[2] for f in feats]
        Ws = [f.shape[3] for f in feats]
        dev, fdt = feats[0].device, feats[0].dtype
        # both filters of every level in one pass over x (arfe_fpn_gate_conv_backward); a missing output
        # gradient is a zero map
        dg = [[(g.to(fdt).contiguous() if g is not None else torch.zeros((B, 1, h, w), dtype=fdt, device=dev))
               for g, h, w in zip(gs, Hs, Ws)] for gs in (grads[:n], grads[n:])]
        need_dx = any(need[1:1 + n])
        mf = torch.channels_last
        dx = [torch.empty(f.shape, dtype=fdt, device=dev, memory_format=mf) for f in feats] if need_dx else None
        f32 = lambda ts: [t.detach().float().contiguous() for t in ts]
        w1f, w2f = f32(w1), f32(w2)
        dw1 = [torch.zeros_like(t) for t in w1f]
        dw2 = [torch.zeros_like(t) for t in w2f]
        db1 = [torch.zeros(1, dtype=torch.float32, device=dev) for _ in range(n)]
        db2 = [torch.zeros(1, dtype=torch.float32, device=dev) for _ in range(n)]
        if B > 0:
            rc = L.lib().arfe_fpn_gate_conv_backward(
                L.ptr_array(feats), L.ptr_array(w1f), L.ptr_array(w2f), L.ptr_array(dg[0]), L.ptr_array(dg[1]),
                L.int_array(Hs), L.int_array(Ws), n, B, C, L.dtype_code(feats[0]), L.ARFE_NHWC,
                L.ptr_array(dx) if need_dx else None, L.ptr_array(dw1), L.ptr_array(db1), L.ptr_array(dw2),
                L.ptr_array(db2), L.stream_ptr(dev))
            L.check(rc, "arfe_fpn_gate_conv_backward")
        elif need_dx:
            dx = [t.zero_() for t in dx]
        pick = lambda ts, off, ref: tuple((t.to(r.dtype) if need[off + i] else None)
                                          for i, (t, r) in enumerate(zip(ts, ref)))
        gx = tuple((dx[l] if need[1 + l] else None) for l in range(n)) if need_dx else (None,) * n
        return (None,) + gx + pick(dw1, 1 + n, w1) + pick(db1, 1 + 2 * n, w1) + pick(dw2, 1 + 3 * n, w2) + \
            pick(db2, 1 + 4 * n, w2)


def fpn_gate_conv(feats, w1, b1, w2, b2):
    """(g1 list, g2 list): raw outputs (bias included, no activation) of the two C -> 1 3x3 gate
    convolutions of every level; feats channels-last; w*: [1, C, 3, 3], b*: [1] per level."""
    n = len(feats)
    out = _FPNGateConvFunction.apply(n, *feats, *w1, *b1, *w2, *b2)
    return list(out[:n]), list(out[n:])


class FPNLink:
    """Couples the gather and the gated residual of ONE neck forward (WFPNDualSpatial.forward):
    x_l feeds both, so autograd would add their two gradients with one more pass over the
    pyramid.  With a link the residual's backward hands its d out_l (= its d x_l) to the
    gather's backward, which runs later (bsf = refine(gather(x)) sits between them) and
    writes d x_l = d out_l + routed gradient once (arfe_fpn_gather_backward_acc).  Only for
    a composition in which the residual's bsf is computed from the gather's output."""

    def __init__(self):
        self.gather_pending = False
        self.douts = None


class _FPNGatherFunction(Function):
    """wfpn_dual_spatial.py:102-113."""

    @staticmethod
    def forward(ctx, refine_level, link, *feats):
        feats, layout, dt = _prep_feats(feats)
        B, C = feats[0].shape[:2]
        Hs = [f.shape[2] for f in feats]
        Ws = [f.shape[3] for f in feats]
        Hr, Wr = Hs[refine_level], Ws[refine_level]
        mf = torch.channels_last if layout == L.ARFE_NHWC else torch.contiguous_format
        out = torch.empty((B, C, Hr, Wr), dtype=feats[0].dtype,
                          device=feats[0].device, memory_format=mf)
        need_grad = any(f.requires_grad for f in feats)
        argmax = None
        if need_grad and refine_level > 0:
            argmax = torch.empty((refine_level, B, C, Hr, Wr), dtype=torch.uint8,
                                 device=out.device)
        if B > 0:
            rc = L.lib().arfe_fpn_gather_forward(
                L.ptr_array(feats), L.int_array(Hs), L.int_array(Ws), len(feats),
                B, C, refine_level, dt, layout, out.data_ptr(),
                argmax.data_ptr() if argmax is not None else None,
                L.stream_ptr(out.device))
            L.check(rc, "arfe_fpn_gather_forward")
        ctx.meta = (refine_level, layout, dt, B, C, Hs, Ws, feats[0].dtype)
        ctx.link = link
        if link is not None:
            link.gather_pending = need_grad and feats[0].dtype == torch.float32 and B > 0
        if argmax is not None:
            ctx.save_for_backward(argmax)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        refine_level, layout, dt, B, C, Hs, Ws, fdtype = ctx.meta
        argmax = ctx.saved_tensors[0] if ctx.saved_tensors else None
        mf = torch.channels_last if layout == L.ARFE_NHWC else torch.contiguous_format
        g = L.as_layout(g.to(fdtype), layout)
        dfeats = [torch.empty((B, C, Hs[l], Ws[l]), dtype=fdtype, device=g.device,
                              memory_format=mf) for l in range(len(Hs))]
        link, addend = ctx.link, None
        if link is not None:
            link.gather_pending = False
            if link.douts is not None:  # the residual's d out_l: folded into the one write of d x_l
                addend, link.douts = L.ptr_array(link.douts), None
        if B > 0:
            rc = L.lib().arfe_fpn_gather_backward_acc(
                g.data_ptr(), argmax.data_ptr() if argmax is not None else None,
                L.int_array(Hs), L.int_array(Ws), len(Hs), B, C, refine_level, dt,
                layout, addend, L.ptr_array(dfeats), L.stream_ptr(g.device))
            L.check(rc, "arfe_fpn_gather_backward")
        return (None, None) + tuple(dfeats)


def fpn_gather(feats, refine_level=2, link=None):
    return _FPNGatherFunction.apply(refine_level, link, *feats)


class _FPNApplyFunction(Function):
    """wfpn_dual_spatial.py:118-135; args: bsf, then L feats, L g1, L g2."""

    @staticmethod
    def forward(ctx, nlev, link, bsf, *tensors):
        ctx.link = link
        feats, g1, g2 = tensors[:nlev], tensors[nlev:2 * nlev], tensors[2 * nlev:]
        feats, layout, dt = _prep_feats(feats)
        L.require_cuda(bsf, *g1, *g2)
        bsf = L.as_layout(bsf.to(feats[0].dtype), layout)
        g1 = [t.to(feats[0].dtype).contiguous() for t in g1]
        g2 = [t.to(feats[0].dtype).contiguous() for t in g2]
        B, C = feats[0].shape[:2]
        Hs = [f.shape[2] for f in feats]
        Ws = [f.shape[3] for f in feats]
        Hr, Wr = bsf.shape[2:]
        outs = [torch.empty_like(f) for f in feats]
        if B > 0:
            rc = L.lib().arfe_fpn_apply_forward(
                L.ptr_array(feats), bsf.data_ptr(), L.ptr_array(g1), L.ptr_array(g2),
                L.int_array(Hs), L.int_array(Ws), nlev, B, C, Hr, Wr, dt, layout,
                L.ptr_array(outs), L.stream_ptr(bsf.device))
            L.check(rc, "arfe_fpn_apply_forward")
        ctx.meta = (nlev, layout, dt, B, C, Hs, Ws, Hr, Wr, feats[0].dtype)
        ctx.save_for_backward(bsf, *g1, *g2)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *douts):
        nlev, layout, dt, B, C, Hs, Ws, Hr, Wr, fdtype = ctx.meta
        saved = ctx.saved_tensors
        bsf, g1, g2 = saved[0], saved[1:1 + nlev], saved[1 + nlev:]
        dev = bsf.device
        mf = torch.channels_last if layout == L.ARFE_NHWC else torch.contiguous_format
        d = []
        for l in range(nlev):
            g = douts[l]
            if g is None:
                g = _zeros((B, C, Hs[l], Ws[l]), fdtype, dev, layout)
            d.append(L.as_layout(g.to(fdtype), layout))
        dbsf = torch.empty((B, C, Hr, Wr), dtype=torch.float32, device=dev,
                           memory_format=mf)
        dg1 = [torch.empty((B, 1, Hs[l], Ws[l]), dtype=torch.float32, device=dev)
               for l in range(nlev)]
        dg2 = [torch.empty_like(t) for t in dg1]
        if B > 0:
            rc = L.lib().arfe_fpn_apply_backward(
                L.ptr_array(d), bsf.data_ptr(), L.ptr_array(list(g1)),
                L.ptr_array(list(g2)), L.int_array(Hs), L.int_array(Ws), nlev, B,
                C, Hr, Wr, dt, layout, dbsf.data_ptr(), L.ptr_array(dg1),
                L.ptr_array(dg2), L.stream_ptr(dev))
            L.check(rc, "arfe_fpn_apply_backward")
        cast = (lambda t: t) if fdtype == torch.float32 else (lambda t: t.to(fdtype))
        # d x_l = d out_l: the residual path is the identity, no copy is made
        dx = tuple(d)
        link = ctx.link
        if link is not None and link.gather_pending and fdtype == torch.float32 and \
                all(ctx.needs_input_grad[3:3 + nlev]):
            # the gather's backward (still to run: bsf came from it) adds them in its own write
            link.douts, dx = list(d), (None,) * nlev
        return (None, None, cast(dbsf)) + dx + tuple(cast(t) for t in dg1) + \
            tuple(cast(t) for t in dg2)


def fpn_apply(feats, bsf, g1, g2, link=None):
    n = len(feats)
    return _FPNApplyFunction.apply(n, link, bsf, *feats, *g1, *g2)


class _NonLocalAttentionFunction(Function):
    """mmdet/ops/non_local.py:65-69 + :98-101: y = softmax(scale * theta_x . phi_x) . g_x as one
    fused tensor-core attention (arfe_nonlocal_attention_forward); the HW x HW weight matrix is
    not materialised.  Backward: recomputed from the saved operands with library GEMMs on bf16
    operands (the forward's arithmetic; a flash-style backward kernel is not built)."""

    @staticmethod
    def forward(ctx, theta, phi, g, scale, nsplit):
        L.require_cuda(theta, phi, g)
        layout = L.layout_of(theta)
        theta, phi, g = (L.as_layout(t, layout) for t in (theta, phi, g))
        dt = L.dtype_code(theta)
        if not (theta.dtype == phi.dtype == g.dtype) or not (theta.shape == phi.shape == g.shape):
            raise ValueError("nonlocal_attention: theta, phi, g must share dtype and shape")
        B, D, H, W = theta.shape
        HW = H * W
        y = torch.empty_like(theta)
        if B > 0 and HW > 0:
            lib = L.lib()
            with torch.cuda.device(theta.device):
                if nsplit is None:
                    nsplit = lib.arfe_nonlocal_default_split(B, HW)
                nbytes = lib.arfe_nonlocal_workspace_bytes(B, HW, D, nsplit)
            if nbytes == 0:
                raise RuntimeError(f"nonlocal_attention: inter_channels must be 64, 128 or 256 (got {D})")
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=theta.device)
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            rc = lib.arfe_nonlocal_attention_forward(
                theta.data_ptr(), phi.data_ptr(), g.data_ptr(), y.data_ptr(), B, HW, D, dt, layout,
                float(scale), nsplit, base, nbytes, L.stream_ptr(theta.device))
            L.check(rc, "arfe_nonlocal_attention_forward")
        ctx.scale = float(scale)
        ctx.save_for_backward(theta, phi, g)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        # Recomputed from the saved operands with library GEMMs in the forward kernel's arithmetic:
        # bf16 operands (theta, phi, g, the weights P, d y, d S), fp32 accumulation and fp32 softmax.
        theta, phi, g = ctx.saved_tensors
        B, D, H, W = theta.shape
        f, h = torch.float32, torch.bfloat16
        flat = lambda t: t.reshape(B, D, -1) if t.is_contiguous() else t.permute(0, 2, 3, 1).reshape(B, -1, D).transpose(1, 2)
        th = flat(theta).transpose(1, 2).to(h)                   # [B, HW, D]
        ph = flat(phi).to(h)                                     # [B, D, HW]
        gx = flat(g).transpose(1, 2).to(h)                       # [B, HW, D]
        dyx = flat(dy).transpose(1, 2).to(h)                     # [B, HW, D]
        HW = H * W
        s = torch.bmm(th, ph, out_dtype=f)                       # [B, HW, HW] logits
        dp = torch.bmm(dyx, gx.transpose(1, 2), out_dtype=f)     # [B, HW, HW]
        pb = torch.empty((B, HW, HW), dtype=h, device=s.device)
        ds = torch.empty_like(pb)
        # softmax, its Jacobian and the two bf16 casts in one pass over S and dP
        rc = L.lib().arfe_nonlocal_backward_rows(s.data_ptr(), dp.data_ptr(), pb.data_ptr(), ds.data_ptr(),
                                                 B * HW, HW, ctx.scale, L.stream_ptr(s.device))
        L.check(rc, "arfe_nonlocal_backward_rows")
        del s, dp
        dg = torch.bmm(pb.transpose(1, 2), dyx, out_dtype=f)     # [B, HW, D]
        dth = torch.bmm(ds, ph.transpose(1, 2), out_dtype=f)     # [B, HW, D]
        dph = torch.bmm(th.transpose(1, 2), ds, out_dtype=f)     # [B, D, HW]

        def back(t, positions_first):
            t = t.transpose(1, 2) if positions_first else t       # -> [B, D, HW]
            t = t.reshape(B, D, H, W).to(theta.dtype)
            return t.contiguous(memory_format=torch.channels_last) if not theta.is_contiguous() else t.contiguous()
        return back(dth, True), back(dph, False), back(dg, True), None, None


def nonlocal_attention(theta, phi, g, scale=1.0, nsplit=None):
    """theta, phi, g: [B, D, H, W] (the outputs of NonLocal2D's 1x1 convolutions, NCHW or
    channels-last); returns y [B, D, H, W] = what the reference reshapes its
    ``matmul(pairwise_weight, g_x)`` into (non_local.py:98-101)."""
    return _NonLocalAttentionFunction.apply(theta, phi, g, scale, nsplit)
