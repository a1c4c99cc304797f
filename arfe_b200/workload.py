"""The benchmark workload of BASELINE.json configs[1]: one training step
(forward + backward) of the region-aware feature path for the images and RoIs
one GPU holds -- Faster R-CNN R50 shapes, 2 images x 512 RoIs per GPU,
800x1344 input, C=256, FPN strides 4..64 -- driven straight through the C ABI
with every buffer pre-allocated (no allocation, no sync in the step).

The convolutions / NonLocal2D / FCs between the kernels stay on PyTorch and are
NOT part of the step (north_star); their outputs (bsf, gate pre-activations,
relu(conv) of the context regions, incoming gradients) are synthetic inputs.

Step (channels-last: 14 launches of our kernels):
  fwd: gather -> apply -> roi_fuse (3 regions: plan, ring kernel, left-out regions) -> gate
  bwd: gate_bwd -> roi_fuse_bwd (bin, pull, inline tiles, flagged regions; reuses the
       forward's plan) -> apply_bwd -> gather_bwd (2 launches)
"""
import math

import torch

from . import _lib as L

STRIDES = (4, 8, 16, 32, 64)
KERNELS_NCHW = ("fpn_gather_fwd", "fpn_apply_fwd", "roi_fuse_fwd", "rff_gate_fwd",
                "rff_gate_bwd", "roi_fuse_bwd", "fpn_apply_bwd", "fpn_gather_bwd")
# channels-last: the AR-FPN backward is one fused pass over the gradient pyramid
KERNELS = KERNELS_NCHW[:6] + ("fpn_bwd",)


def pyramid_shapes(img_h=800, img_w=1344, strides=STRIDES):
    shapes, h, w, s = [], img_h, img_w, 1
    for st in strides:
        while s < st:
            h, w, s = (h - 1) // 2 + 1, (w - 1) // 2 + 1, s * 2
        shapes.append((h, w))
    return shapes


def synthetic_rois(K, img_w=1344, img_h=800, batch=1, seed=0, smin=16.0, smax=600.0,
                   order="image_major"):
    """centre uniform, sqrt(area) log-uniform in
    [16,600] px, aspect log-uniform in [0.5,2], clipped (same generator as
    oracle.synthetic_rois).  order="image_major": per-image blocks, the order
    bbox2roi builds (mmdet/core/bbox/transforms.py:51-59: one block of
    (img_id, x1, y1, x2, y2) rows per image, torch.cat); "interleaved": RoI i
    belongs to image i % batch (kept as a correctness case)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(K, 4, generator=g)
    cx, cy = u[:, 0] * img_w, u[:, 1] * img_h
    s = torch.exp(math.log(smin) + u[:, 2] * (math.log(smax) - math.log(smin)))
    ar = torch.exp(math.log(0.5) + u[:, 3] * (math.log(2.0) - math.log(0.5)))
    w, h = s * torch.sqrt(ar), s / torch.sqrt(ar)
    x1 = (cx - w / 2).clamp(0, img_w - 1)
    y1 = (cy - h / 2).clamp(0, img_h - 1)
    x2 = torch.max((cx + w / 2).clamp(0, img_w - 1), x1 + 1.0)
    y2 = torch.max((cy + h / 2).clamp(0, img_h - 1), y1 + 1.0)
    if order == "image_major":
        b = ((torch.arange(K) * batch) // max(K, 1)).float()
    elif order == "interleaved":
        b = (torch.arange(K) % batch).float()
    else:
        raise ValueError(f"unknown RoI order {order!r}")
    return torch.stack([b, x1, y1, x2, y2], dim=1).float().contiguous()


def host_inputs(batch=2, rois_per_img=512, channels=256, img_h=800, img_w=1344,
                dtype=torch.float32, seed=0, pin=False, channels_last=False, strides=STRIDES,
                out_size=7, roi_order="image_major", smin=16.0, smax=600.0, device=None):
    """Synthetic inputs of one step on the HOST (optionally pinned; optionally
    stored in torch.channels_last memory format -- same logical tensors).
    device: generate the dense tensors there instead (large sweeps: no host copy
    to make; the RoIs always come from the seeded host generator)."""
    shapes = pyramid_shapes(img_h, img_w, strides)
    g = torch.Generator(device=device).manual_seed(seed) if device is not None else \
        torch.Generator().manual_seed(seed)
    K = batch * rois_per_img
    hr, wr = shapes[2]

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=device)
    t = {}
    t["x"] = [randn(batch, channels, h, w).to(dtype) for h, w in shapes]
    t["bsf"] = randn(batch, channels, hr, wr).to(dtype)
    t["g1"] = [randn(batch, 1, h, w).to(dtype) for h, w in shapes]
    t["g2"] = [randn(batch, 1, h, w).to(dtype) for h, w in shapes]
    t["rois"] = synthetic_rois(K, img_w, img_h, batch, seed, smin=smin, smax=smax, order=roi_order)
    P = out_size
    t["a"] = randn(K, channels, P, P).relu().to(dtype)
    t["b"] = randn(K, channels, P, P).relu().to(dtype)
    t["gz"] = randn(K, channels, P, P).to(dtype)      # dL/d(gate out)
    # dL/d(lw), dL/d(lh): what the backward of the two context-branch convs (PyTorch) hands back
    t["glw"] = randn(K, channels, P, P).to(dtype)
    t["glh"] = randn(K, channels, P, P).to(dtype)
    t["gbsf"] = randn(batch, channels, hr, wr).to(dtype)  # dL/d(gathered)
    if channels_last:
        def cl(v):
            if isinstance(v, list):
                return [cl(e) for e in v]
            return v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v
        t = {k: cl(v) for k, v in t.items()}
    if pin:
        def p(v):
            return [p(e) for e in v] if isinstance(v, list) else v.pin_memory()
        t = {k: p(v) for k, v in t.items()}
    t["shapes"] = shapes
    t["strides"] = tuple(strides)
    t["out_size"] = out_size
    return t


class TrainStep:
    """Pre-allocated device state + the step through the C ABI.

    layout = channels-last (default): the fast path -- 128-bit gathers forward,
    atomic-free pull backward, no zero-fill.  layout = NCHW: the reference's
    memory layout through the compatibility kernels.
    """

    def __init__(self, host, device, regions=3, channels_last=True, split=None, roi_levels=4):
        """split (channels-last only, default on): the region features and their
        gradients are separate tensors (ori | lw | lh) instead of one concatenated
        tensor -- what the head's convolutions produce / consume -- so the step
        needs no torch copies between our kernels.
        roi_levels: pyramid levels the RoI extractor reads -- the reference's
        bbox_roi_extractor has featmap_strides=[4, 8, 16, 32]
        (configs/_base_/models/faster_rcnn_r50_fpn.py:44) and is handed
        x[:num_inputs] (standard_roi_head.py:140); the 5th level only feeds the RPN."""
        self.dev = device
        self.lib = L.lib()
        x = host["x"]
        self.cl = bool(channels_last)
        self.split = self.cl if split is None else (bool(split) and self.cl)
        self.overlap_plan = self.cl  # step(): plan + tile bins on a second stream (plan_async)
        self.layout = L.ARFE_NHWC if self.cl else L.ARFE_NCHW
        mf = torch.channels_last if self.cl else torch.contiguous_format
        self.dtype = x[0].dtype
        self.dt = L.dtype_code(x[0])
        self.B, self.C = x[0].shape[:2]
        self.shapes = host["shapes"]
        self.nlev = len(x)
        self.rlev = min(int(roi_levels), self.nlev)
        self.K = host["rois"].shape[0]
        self.R = regions
        self.P = host.get("out_size", 7)
        self.PP = self.P * self.P
        self.strides = host.get("strides", STRIDES)

        def d(v):
            if isinstance(v, list):
                return [d(e) for e in v]
            v = v.to(device)
            return v.contiguous(memory_format=mf) if v.dim() == 4 else v
        self.x, self.bsf, self.g1, self.g2 = d(host["x"]), d(host["bsf"]), d(host["g1"]), d(host["g2"])
        self.rois, self.a, self.b, self.gz, self.gbsf = (d(host[k]) for k in ("rois", "a", "b", "gz", "gbsf"))
        self.glw, self.glh = d(host["glw"]), d(host["glh"])
        B, C, K, R = self.B, self.C, self.K, self.R
        hr, wr = self.shapes[2]

        def e(*s, dtype=self.dtype):
            if len(s) == 4:
                return torch.empty(s, dtype=dtype, device=device, memory_format=mf)
            return torch.empty(s, dtype=dtype, device=device)
        self.gathered = e(B, C, hr, wr)
        self.argmax = torch.empty((2, B, C, hr, wr), dtype=torch.uint8, device=device)
        self.y = [e(B, C, h, w) for h, w in self.shapes]
        P = self.P
        self.F = e(K, R * C, P, P)
        self.z = e(K, C, P, P)
        self.dF = e(K, R * C, P, P)            # [d_ori | d_lw | d_lh]
        self.d_ori = e(K, C, P, P)
        self.d_ab = e(K, C, P, P)
        if self.split:
            self.Fr = [e(K, C, P, P) for _ in range(R)]
            self.p_Fr = L.ptr_array(self.Fr)
            # d_ori from the gate backward; d_lw / d_lh stand in for the backward of the two
            # context-branch convs (PyTorch, out of the path): tensors of their own
            self.p_dFr = L.ptr_array([self.d_ori, self.glw, self.glh][:R])
        self.dy = [e(B, C, h, w, dtype=torch.float32) for h, w in self.shapes]
        for t in self.dy[self.rlev:]:
            t.zero_()  # levels the RoI head does not read: no gradient from this path, never rewritten
        self.dy_in = self.dy if self.dtype == torch.float32 else [e(B, C, h, w) for h, w in self.shapes]
        self.dbsf = e(B, C, hr, wr, dtype=torch.float32)
        self.dg1 = [e(B, 1, h, w, dtype=torch.float32) for h, w in self.shapes]
        self.dg2 = [e(B, 1, h, w, dtype=torch.float32) for h, w in self.shapes]
        self.dx = [e(B, C, h, w) for h, w in self.shapes]
        self.ws_bytes = self.lib.arfe_roi_plan_bytes(
            K, R, self.rlev, B, L.int_array([s[0] for s in self.shapes]),
            L.int_array([s[1] for s in self.shapes]))
        self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=device)
        self.ws_ptr = (self.ws.data_ptr() + 255) // 256 * 256
        # C arrays built once
        self.H = L.int_array([s[0] for s in self.shapes])
        self.W = L.int_array([s[1] for s in self.shapes])
        self.scales = L.float_array([1.0 / s for s in self.strides[:self.nlev]])
        self.p_x, self.p_y = L.ptr_array(self.x), L.ptr_array(self.y)
        self.p_g1, self.p_g2 = L.ptr_array(self.g1), L.ptr_array(self.g2)
        self.p_dy, self.p_dy_in = L.ptr_array(self.dy), L.ptr_array(self.dy_in)
        self.p_dg1, self.p_dg2 = L.ptr_array(self.dg1), L.ptr_array(self.dg2)
        self.p_dx = L.ptr_array(self.dx)
        # the gate sees [rows][n] with `ori` strided inside the concatenated tensor
        self.gate_in, self.gate_dout = self.F, self.dF
        if self.split:  # memory [K*49][C]: `ori` and d_ori are tensors of their own
            self.gate_rows, self.gate_n, self.gate_stride = K * self.PP, C, C
            self.gate_in, self.gate_dout = self.Fr[0], self.d_ori
        elif self.cl:   # memory [K*49][R*C]: a row is one bin of one RoI
            self.gate_rows, self.gate_n, self.gate_stride = K * self.PP, C, R * C
        else:         # memory [K][R*C*49]: a row is one RoI
            self.gate_rows, self.gate_n, self.gate_stride = K, C * self.PP, R * C * self.PP
        self.stream = L.stream_ptr(device)
        self.kernels = KERNELS if self.cl else KERNELS_NCHW

    # -- the eight ops; each returns the C return code ----------------------
    def fpn_gather_fwd(self):
        return self.lib.arfe_fpn_gather_forward(
            self.p_x, self.H, self.W, self.nlev, self.B, self.C, 2, self.dt, self.layout,
            self.gathered.data_ptr(), self.argmax.data_ptr(), self.stream)

    def fpn_apply_fwd(self):
        hr, wr = self.shapes[2]
        return self.lib.arfe_fpn_apply_forward(
            self.p_x, self.bsf.data_ptr(), self.p_g1, self.p_g2, self.H, self.W, self.nlev,
            self.B, self.C, hr, wr, self.dt, self.layout, self.p_y, self.stream)

    def _geo(self):
        return (self.H, self.W, self.scales, self.rlev, self.B, self.C, self.rois.data_ptr(), self.K,
                self.R, 1.0, self.P, self.P, 0, 56.0, self.dt)

    def _side(self):
        if getattr(self, "side", None) is None:
            self.side = torch.cuda.Stream(self.dev)
            self.ev_fork, self.ev_plan, self.ev_bin = (torch.cuda.Event() for _ in range(3))
        return self.side

    def plan_async(self, bins=True):
        """The RoI plan depends only on the RoIs: build it on a second stream, under the AR-FPN
        forward kernels (latency-bound table walking next to bandwidth-bound streaming).
        bins: the backward's tile bins follow -- bins_async(), issued after the forward."""
        side = self._side()
        # fork: everything before this step (the previous step's backward still reads the
        # workspace) precedes the rebuild; the join is the wait on ev_plan
        self.ev_fork.record(torch.cuda.current_stream(self.dev))
        side.wait_event(self.ev_fork)
        L.check(self.lib.arfe_roi_plan_build(*self._geo(), self.ws_ptr, self.ws_bytes, side.cuda_stream),
                "arfe_roi_plan_build")
        self.ev_plan.record(side)
        self.async_plan, self.async_bins, self.want_bins_now = 1, 0, bool(bins)

    def bins_async(self):
        """The pull backward's tile bins depend only on the plan.  They are forked AFTER the
        forward's launch: next to the ring forward (another latency-bound kernel) the binning
        kernel cost it ~20 us; next to the streaming gate kernels that follow it is free."""
        side = self._side()
        self.ev_fork.record(torch.cuda.current_stream(self.dev))
        side.wait_event(self.ev_fork)
        L.check(self.lib.arfe_roi_pull_bin(*self._geo(), int(self.split), self.ws_ptr, self.ws_bytes,
                                           side.cuda_stream), "arfe_roi_pull_bin")
        self.ev_bin.record(side)
        self.async_bins, self.want_bins_now = 1, False

    def roi_fuse_fwd(self):
        ready = 0
        if self.cl and getattr(self, "async_plan", 0):
            torch.cuda.current_stream(self.dev).wait_event(self.ev_plan)
            ready, self.async_plan = 1, 0
        if self.split:
            self.planned = 1
            return self.lib.arfe_roi_fuse_forward_plan_split(
                self.p_y, self.H, self.W, self.scales, self.rlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.p_Fr, self.ws_ptr, self.ws_bytes, ready, self.stream)
        if self.cl:
            # channels-last: plan + ring kernel; the plan stays in the workspace for the backward
            self.planned = 1
            return self.lib.arfe_roi_fuse_forward_plan(
                self.p_y, self.H, self.W, self.scales, self.rlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.F.data_ptr(), None, None, self.ws_ptr, self.ws_bytes, ready, self.stream)
        return self.lib.arfe_roi_fuse_forward(
            self.p_y, self.H, self.W, self.scales, self.rlev, self.B, self.C,
            self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt, self.layout,
            self.layout, self.F.data_ptr(), None, None, self.stream)

    def rff_gate_fwd(self):
        return self.lib.arfe_rff_gate_forward(
            self.gate_in.data_ptr(), self.gate_stride, self.a.data_ptr(), self.b.data_ptr(),
            self.z.data_ptr(), self.gate_rows, self.gate_n, self.dt, self.stream)

    def rff_gate_bwd(self):
        return self.lib.arfe_rff_gate_backward(
            self.gz.data_ptr(), self.gate_in.data_ptr(), self.gate_stride, self.a.data_ptr(),
            self.b.data_ptr(), self.gate_dout.data_ptr(), self.gate_stride, self.d_ab.data_ptr(),
            self.gate_rows, self.gate_n, self.dt, self.stream)

    def roi_fuse_bwd(self):
        ready = getattr(self, "planned", 0)
        if self.cl and getattr(self, "async_bins", 0):
            torch.cuda.current_stream(self.dev).wait_event(self.ev_bin)
            ready, self.async_bins = 2, 0
        if self.split:
            rc = self.lib.arfe_roi_fuse_backward_pull_split(
                self.p_dFr, self.H, self.W, self.scales, self.rlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.p_dy, self.ws_ptr, self.ws_bytes, ready, self.stream)
        elif self.cl:
            rc = self.lib.arfe_roi_fuse_backward_pull(
                self.dF.data_ptr(), self.H, self.W, self.scales, self.rlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.p_dy, self.ws_ptr, self.ws_bytes, ready, self.stream)
        if self.cl:
            return rc
        return self.lib.arfe_roi_fuse_backward(
            self.dF.data_ptr(), L.ARFE_NCHW, self.H, self.W, self.scales, self.rlev, self.B,
            self.C, self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
            self.layout, self.p_dy, self.stream)

    def fpn_apply_bwd(self):
        hr, wr = self.shapes[2]
        return self.lib.arfe_fpn_apply_backward(
            self.p_dy_in, self.bsf.data_ptr(), self.p_g1, self.p_g2, self.H, self.W, self.nlev,
            self.B, self.C, hr, wr, self.dt, self.layout, self.dbsf.data_ptr(), self.p_dg1,
            self.p_dg2, self.stream)

    def fpn_gather_bwd(self):
        # d x_l = d out_l (the residual's identity, = dy_l written by roi_fuse_bwd) + the gather's
        # routed gradient, in the one write of dx_l (wfpn_dual_spatial.py:113,135 under autograd)
        return self.lib.arfe_fpn_gather_backward_acc(
            self.gbsf.data_ptr(), self.argmax.data_ptr(), self.H, self.W, self.nlev, self.B,
            self.C, 2, self.dt, self.layout, self.p_dy, self.p_dx, self.stream)

    def fpn_bwd(self):
        """apply backward + gather backward fused (channels-last): dbsf, dg1, dg2 and
        d x_l = d out_l + the gather's routed gradient, d out (fp32, straight from the RoI
        backward) read once."""
        return self.lib.arfe_fpn_backward_fused(
            self.p_dy, 1, self.bsf.data_ptr(), self.p_g1, self.p_g2, self.gbsf.data_ptr(),
            self.argmax.data_ptr(), self.H, self.W, self.nlev, self.B, self.C, 2, self.dt, self.layout,
            self.dbsf.data_ptr(), self.p_dg1, self.p_dg2, self.p_dx, self.stream)

    def glue_before_roi_bwd(self):
        """torch plumbing between our kernels: assemble dF (stand-in for the
        conv backward of the two context branches, which stays on PyTorch) and,
        on the NCHW path only, zero the accumulators (the reference's
        at::zeros, roi_align_kernel_v2.cu:325-326; the pull kernel needs none)."""
        if self.split:
            return  # the pull kernel reads d_ori / d_ab where they are
        C = self.C  # d_ori was written in place into dF[:, :C] by the gate backward
        for r, t in zip(range(1, self.R), (self.glw, self.glh)):
            self.dF[:, r * C:(r + 1) * C].copy_(t)
        if not self.cl:
            for t in self.dy[:self.rlev]:
                t.zero_()

    def glue_before_apply_bwd(self):
        if self.dy_in is not self.dy:
            for a, b in zip(self.dy_in, self.dy):
                a.copy_(b)

    _GLUE = {"roi_fuse_bwd": "glue_before_roi_bwd", "fpn_apply_bwd": "glue_before_apply_bwd"}

    def run_ops(self, names, timer=None):
        """The named ops in order (with the torch glue that belongs in front of them);
        `timer(name, fn)` wraps each of our launches."""
        run = timer or (lambda name, fn: L.check(fn(), name))
        if self.cl and self.overlap_plan and "roi_fuse_fwd" in names:
            self.plan_async(bins="roi_fuse_bwd" in names or getattr(self, "want_bins", False))
        for n in names:
            glue = self._GLUE.get(n)
            if glue:
                getattr(self, glue)()
            run(n, getattr(self, n))
            if n == "roi_fuse_fwd" and getattr(self, "want_bins_now", False):
                self.bins_async()

    def step(self, timer=None):
        """One training step. `timer(name, fn)` wraps each of our launches; without a
        timer a captured step (capture()) is replayed as one CUDA graph."""
        if timer is None and getattr(self, "graph", None) is not None:
            self.graph.replay()
            return
        self.run_ops(self.kernels, timer)

    def capture(self):
        """Capture the step (the same launches, second stream included) into a CUDA graph:
        the 14 kernels are 35-170 us each, so launch gaps are a visible share of the step."""
        for _ in range(3):
            self.step()  # kernel attributes set, second stream and events created
        torch.cuda.synchronize(self.dev)
        graph, cap, saved = torch.cuda.CUDAGraph(), torch.cuda.Stream(self.dev), self.stream
        try:
            with torch.cuda.graph(graph, stream=cap):
                self.stream = cap.cuda_stream
                self.step()
        finally:
            self.stream = saved
        self.graph = graph

    def launches_per_step(self):
        """Kernels of libarfe_b200.so per step: gather 1, apply 1, roi fwd (plan +
        ring + left-out regions = 3 | 1), gate 2, roi bwd (bin + pull + inline tiles +
        flagged-region fallback = 4, the plan is the forward's | atomic 1),
        apply bwd 1 | 2, gather bwd 2 (pooled levels + small levels; one merged launch was
        measured slower: 49 us vs 34 + 11 us, a third wave of CTAs) | 2.  (Channels-last:
        plan and bin run on the second stream of plan_async; same count.)"""
        return (1 + 1 + 3 + 2 + 4 + 2) if self.cl else (1 + 1 + 1 + 2 + 1 + 2 + 2)

    # -- algorithmic bytes per launch ------------------
    def algorithmic_bytes(self):
        e = 4 if self.dtype == torch.float32 else 2
        B, C, K, R = self.B, self.C, self.K, self.R
        P = sum(h * w for h, w in self.shapes)
        Pr = sum(h * w for h, w in self.shapes[:self.rlev])  # levels the RoI extractor reads
        hr, wr = self.shapes[2]
        pyr = B * C * P * e
        ref = B * C * hr * wr
        out_roi = K * R * C * self.PP * e
        n_gate = K * C * self.PP * e
        return {
            "fpn_gather_fwd": pyr + ref * e + 2 * ref,                       # + uint8 argmax (2 levels)
            "fpn_apply_fwd": 2 * pyr + ref * e + 2 * B * P * e,
            "roi_fuse_fwd": out_roi + B * C * Pr * e + 20 * K,
            "rff_gate_fwd": 4 * n_gate,
            "rff_gate_bwd": 6 * n_gate,
            "roi_fuse_bwd": out_roi + B * C * Pr * 4 + 20 * K,               # fp32 accumulators
            "fpn_apply_bwd": pyr + ref * e + 2 * B * P * e + ref * 4 + 2 * B * P * 4,
            "fpn_gather_bwd": ref * e + 2 * ref + pyr + B * C * P * 4,          # + the fp32 addend (dy)
            # fused: d out (fp32) once, d x once, bsf + d gathered + argmax + dbsf, gate maps in and out
            "fpn_bwd": B * C * P * 4 + pyr + 2 * ref * e + 2 * ref + ref * 4 + 2 * B * P * e + 2 * B * P * 4,
        }


# ---------------------------------------------------------------------------
# The other BASELINE.json configurations, as schedules over TrainStep states
# ---------------------------------------------------------------------------
FWD_OPS = ("fpn_gather_fwd", "fpn_apply_fwd", "roi_fuse_fwd", "rff_gate_fwd")
ROI_FWD = ("roi_fuse_fwd", "rff_gate_fwd")
ROI_BWD = ("rff_gate_bwd", "roi_fuse_bwd")
FPN_BWD = ("fpn_bwd",)
RETINA_STRIDES = (8, 16, 32, 64, 128)


class Case:
    """An ordered schedule [(prefix, TrainStep, op names)] = one step of a configuration."""

    def __init__(self, workload, images, parts, dtype):
        self.workload, self.images, self.parts, self.dtype = workload, images, parts, dtype

    def step(self, timer=None):
        for prefix, st, names in self.parts:
            t = None if timer is None else (lambda n, fn, p=prefix: timer(p + n, fn))
            st.run_ops(names, t)

    def op_names(self):
        return [p + n for p, _, names in self.parts for n in names]

    def algorithmic_bytes(self):
        out = {}
        for prefix, st, names in self.parts:
            alg = st.algorithmic_bytes()
            for n in names:
                out[prefix + n] = alg[n]
        return out

    def launches_per_step(self):
        cl = {"fpn_gather_fwd": 1, "fpn_apply_fwd": 1, "roi_fuse_fwd": 3, "rff_gate_fwd": 1,
              "rff_gate_bwd": 1, "roi_fuse_bwd": 4, "fpn_apply_bwd": 1, "fpn_gather_bwd": 2, "fpn_bwd": 2}
        nchw = dict(cl, roi_fuse_fwd=1, roi_fuse_bwd=1, fpn_apply_bwd=2)
        return sum((cl if st.cl else nchw)[n] for _, st, names in self.parts for n in names)


def _share_pyramid(st, src):
    """st's RoI ops read src's aggregated pyramid (one neck, several RoI consumers)."""
    st.y, st.p_y = src.y, src.p_y
    return st


def make_case(config, dev, seed=0, rois_per_img=None):
    """BASELINE.json configs[config] on one GPU (channels-last fast path, dense inputs
    generated on the device)."""
    f32, bf16 = torch.float32, torch.bfloat16
    if config == 0:
        K = rois_per_img or 1000
        st = TrainStep(host_inputs(1, K, 256, seed=seed, channels_last=True, device=dev), dev)
        return Case(f"configs[0]: Faster R-CNN R50 + AR-FPN + AR-RFF inference (forward of the fused path), "
                    f"1 image 800x1344, {K} RoIs, C=256, fp32", 1, [("", st, FWD_OPS)], "f32")
    if config == 1:
        K = rois_per_img or 512
        st = TrainStep(host_inputs(2, K, 256, seed=seed, channels_last=True, device=dev), dev)
        return Case(f"configs[1]: training step, 2 img x {K} RoIs", 2, [("", st, KERNELS)], "f32")
    if config == 2:
        st = TrainStep(host_inputs(8, 8, 256, seed=seed, channels_last=True, strides=RETINA_STRIDES,
                                   device=dev), dev)
        return Case("configs[2]: RetinaNet R50 + AR-FPN neck-only aggregation (gather + gated residual, "
                    "forward), batch 8, 800x1344, strides 8-128, C=256, fp32", 8,
                    [("", st, FWD_OPS[:2])], "f32")
    if config == 3:
        K = rois_per_img or 512
        st = TrainStep(host_inputs(2, K, 256, dtype=bf16, seed=seed, channels_last=True, device=dev), dev)
        km = max(K // 4, 1)  # mask head: the positive quarter of the sampled RoIs
        mask = TrainStep(host_inputs(2, km, 256, dtype=bf16, seed=seed + 1, channels_last=True,
                                     out_size=14, device=dev), dev, regions=1)
        mask.d_ori.normal_()  # incoming gradient of the 14x14 mask features
        _share_pyramid(mask, st)
        st.want_bins = mask.want_bins = True  # the backward follows in a later part of the schedule
        return Case(f"configs[3]: Mask R-CNN R50 + AR-FPN + AR-RFF training step, bf16 I/O, 2 img x {K} RoIs "
                    f"(bbox 7x7 x 3 regions) + 2 x {km} RoIs (mask 14x14 x 1 region)", 2,
                    [("", st, KERNELS[:4]), ("mask.", mask, ("roi_fuse_fwd",)),
                     ("mask.", mask, ("roi_fuse_bwd",)), ("", st, KERNELS[4:])], "bf16")
    if config == 4:
        K = rois_per_img or 512
        stages = [TrainStep(host_inputs(1, K, 256, seed=seed + i, channels_last=True, device=dev), dev)
                  for i in range(3)]
        for st in stages:
            st.want_bins = True
        for st in stages[1:]:
            _share_pyramid(st, stages[0])
        parts = [("", stages[0], FWD_OPS[:2])]
        parts += [(f"s{i}.", st, ROI_FWD) for i, st in enumerate(stages)]
        parts += [(f"s{i}.", st, ROI_BWD) for i, st in reversed(list(enumerate(stages)))]
        parts += [("", stages[0], FPN_BWD)]
        return Case(f"configs[4]: Cascade R-CNN + AR-RFF, three-stage RoI fusion training step, 1 image x {K} "
                    f"RoIs per stage, C=256, fp32", 1, parts, "f32")
    raise ValueError(f"unknown config {config}")
