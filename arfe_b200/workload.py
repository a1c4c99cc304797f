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
KERNELS = ("fpn_gather_fwd", "fpn_apply_fwd", "roi_fuse_fwd", "rff_gate_fwd",
           "rff_gate_bwd", "roi_fuse_bwd", "fpn_apply_bwd", "fpn_gather_bwd")


def pyramid_shapes(img_h=800, img_w=1344, strides=STRIDES):
    shapes, h, w, s = [], img_h, img_w, 1
    for st in strides:
        while s < st:
            h, w, s = (h - 1) // 2 + 1, (w - 1) // 2 + 1, s * 2
        shapes.append((h, w))
    return shapes


def synthetic_rois(K, img_w=1344, img_h=800, batch=1, seed=0, smin=16.0, smax=600.0):
    """centre uniform, sqrt(area) log-uniform in
    [16,600] px, aspect log-uniform in [0.5,2], clipped; RoI i belongs to
    image i % batch (same generator as oracle.synthetic_rois)."""
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
    b = (torch.arange(K) % batch).float()
    return torch.stack([b, x1, y1, x2, y2], dim=1).float().contiguous()


def host_inputs(batch=2, rois_per_img=512, channels=256, img_h=800, img_w=1344,
                dtype=torch.float32, seed=0, pin=False, channels_last=False, strides=STRIDES,
                out_size=7):
    """Synthetic inputs of one step on the HOST (optionally pinned; optionally
    stored in torch.channels_last memory format -- same logical tensors)."""
    shapes = pyramid_shapes(img_h, img_w, strides)
    g = torch.Generator().manual_seed(seed)
    K = batch * rois_per_img
    hr, wr = shapes[2]
    t = {}
    t["x"] = [torch.randn(batch, channels, h, w, generator=g).to(dtype) for h, w in shapes]
    t["bsf"] = torch.randn(batch, channels, hr, wr, generator=g).to(dtype)
    t["g1"] = [torch.randn(batch, 1, h, w, generator=g).to(dtype) for h, w in shapes]
    t["g2"] = [torch.randn(batch, 1, h, w, generator=g).to(dtype) for h, w in shapes]
    t["rois"] = synthetic_rois(K, img_w, img_h, batch, seed)
    P = out_size
    t["a"] = torch.randn(K, channels, P, P, generator=g).relu().to(dtype)
    t["b"] = torch.randn(K, channels, P, P, generator=g).relu().to(dtype)
    t["gz"] = torch.randn(K, channels, P, P, generator=g).to(dtype)      # dL/d(gate out)
    t["gbsf"] = torch.randn(batch, channels, hr, wr, generator=g).to(dtype)  # dL/d(gathered)
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

    def __init__(self, host, device, regions=3, channels_last=True, split=None):
        """split (channels-last only, default on): the region features and their
        gradients are separate tensors (ori | lw | lh) instead of one concatenated
        tensor -- what the head's convolutions produce / consume -- so the step
        needs no torch copies between our kernels."""
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
            # stand-ins for the conv backward of the two context branches: d_ab for both
            self.p_dFr = L.ptr_array([self.d_ori] + [self.d_ab] * (R - 1))
        self.dy = [e(B, C, h, w, dtype=torch.float32) for h, w in self.shapes]
        self.dy_in = self.dy if self.dtype == torch.float32 else [e(B, C, h, w) for h, w in self.shapes]
        self.dbsf = e(B, C, hr, wr, dtype=torch.float32)
        self.dg1 = [e(B, 1, h, w, dtype=torch.float32) for h, w in self.shapes]
        self.dg2 = [e(B, 1, h, w, dtype=torch.float32) for h, w in self.shapes]
        self.dx = [e(B, C, h, w) for h, w in self.shapes]
        self.ws_bytes = self.lib.arfe_roi_plan_bytes(
            K, R, self.nlev, B, L.int_array([s[0] for s in self.shapes]),
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

    def plan_async(self):
        """The RoI plan depends only on the RoIs, the backward's tile bins only on
        the plan: build both on a second stream, under the AR-FPN kernels of the
        forward (latency-bound table walking next to bandwidth-bound streaming)."""
        if getattr(self, "side", None) is None:
            self.side = torch.cuda.Stream(self.dev)
            self.ev_fork, self.ev_plan, self.ev_bin = (torch.cuda.Event() for _ in range(3))
        # fork: everything before this step (the previous step's backward still reads the
        # workspace) precedes the rebuild; the joins are the waits on ev_plan / ev_bin
        self.ev_fork.record(torch.cuda.current_stream(self.dev))
        self.side.wait_event(self.ev_fork)
        geo = (self.H, self.W, self.scales, self.nlev, self.B, self.C, self.rois.data_ptr(), self.K,
               self.R, 1.0, self.P, self.P, 0, 56.0, self.dt)
        tail = (self.ws_ptr, self.ws_bytes, self.side.cuda_stream)
        L.check(self.lib.arfe_roi_plan_build(*geo, *tail), "arfe_roi_plan_build")
        self.ev_plan.record(self.side)
        L.check(self.lib.arfe_roi_pull_bin(*geo, int(self.split), *tail), "arfe_roi_pull_bin")
        self.ev_bin.record(self.side)
        self.async_plan, self.async_bins = 1, 1

    def roi_fuse_fwd(self):
        ready = 0
        if self.cl and getattr(self, "async_plan", 0):
            torch.cuda.current_stream(self.dev).wait_event(self.ev_plan)
            ready, self.async_plan = 1, 0
        if self.split:
            self.planned = 1
            return self.lib.arfe_roi_fuse_forward_plan_split(
                self.p_y, self.H, self.W, self.scales, self.nlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.p_Fr, self.ws_ptr, self.ws_bytes, ready, self.stream)
        if self.cl:
            # channels-last: plan + ring kernel; the plan stays in the workspace for the backward
            self.planned = 1
            return self.lib.arfe_roi_fuse_forward_plan(
                self.p_y, self.H, self.W, self.scales, self.nlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.F.data_ptr(), None, None, self.ws_ptr, self.ws_bytes, ready, self.stream)
        return self.lib.arfe_roi_fuse_forward(
            self.p_y, self.H, self.W, self.scales, self.nlev, self.B, self.C,
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
                self.p_dFr, self.H, self.W, self.scales, self.nlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.p_dy, self.ws_ptr, self.ws_bytes, ready, self.stream)
        elif self.cl:
            rc = self.lib.arfe_roi_fuse_backward_pull(
                self.dF.data_ptr(), self.H, self.W, self.scales, self.nlev, self.B, self.C,
                self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
                self.p_dy, self.ws_ptr, self.ws_bytes, ready, self.stream)
        if self.cl:
            return rc
        return self.lib.arfe_roi_fuse_backward(
            self.dF.data_ptr(), L.ARFE_NCHW, self.H, self.W, self.scales, self.nlev, self.B,
            self.C, self.rois.data_ptr(), self.K, self.R, 1.0, self.P, self.P, 0, 56.0, self.dt,
            self.layout, self.p_dy, self.stream)

    def fpn_apply_bwd(self):
        hr, wr = self.shapes[2]
        return self.lib.arfe_fpn_apply_backward(
            self.p_dy_in, self.bsf.data_ptr(), self.p_g1, self.p_g2, self.H, self.W, self.nlev,
            self.B, self.C, hr, wr, self.dt, self.layout, self.dbsf.data_ptr(), self.p_dg1,
            self.p_dg2, self.stream)

    def fpn_gather_bwd(self):
        return self.lib.arfe_fpn_gather_backward(
            self.gbsf.data_ptr(), self.argmax.data_ptr(), self.H, self.W, self.nlev, self.B,
            self.C, 2, self.dt, self.layout, self.p_dx, self.stream)

    def glue_before_roi_bwd(self):
        """torch plumbing between our kernels: assemble dF (stand-in for the
        conv backward of the two context branches, which stays on PyTorch) and,
        on the NCHW path only, zero the accumulators (the reference's
        at::zeros, roi_align_kernel_v2.cu:325-326; the pull kernel needs none)."""
        if self.split:
            return  # the pull kernel reads d_ori / d_ab where they are
        C = self.C  # d_ori was written in place into dF[:, :C] by the gate backward
        for r in range(1, self.R):
            self.dF[:, r * C:(r + 1) * C].copy_(self.d_ab)
        if not self.cl:
            for t in self.dy:
                t.zero_()

    def glue_before_apply_bwd(self):
        if self.dy_in is not self.dy:
            for a, b in zip(self.dy_in, self.dy):
                a.copy_(b)

    def step(self, timer=None):
        """One training step. `timer(name, fn)` wraps each of our launches; without a
        timer a captured step (capture()) is replayed as one CUDA graph."""
        if timer is None and getattr(self, "graph", None) is not None:
            self.graph.replay()
            return
        run = timer or (lambda name, fn: L.check(fn(), name))
        if self.cl and self.overlap_plan:
            self.plan_async()
        run("fpn_gather_fwd", self.fpn_gather_fwd)
        run("fpn_apply_fwd", self.fpn_apply_fwd)
        run("roi_fuse_fwd", self.roi_fuse_fwd)
        run("rff_gate_fwd", self.rff_gate_fwd)
        run("rff_gate_bwd", self.rff_gate_bwd)
        self.glue_before_roi_bwd()
        run("roi_fuse_bwd", self.roi_fuse_bwd)
        self.glue_before_apply_bwd()
        run("fpn_apply_bwd", self.fpn_apply_bwd)
        run("fpn_gather_bwd", self.fpn_gather_bwd)

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
        return (1 + 1 + 3 + 2 + 4 + 1 + 2) if self.cl else (1 + 1 + 1 + 2 + 1 + 2 + 2)

    # -- algorithmic bytes per launch ------------------
    def algorithmic_bytes(self):
        e = 4 if self.dtype == torch.float32 else 2
        B, C, K, R = self.B, self.C, self.K, self.R
        P = sum(h * w for h, w in self.shapes)
        hr, wr = self.shapes[2]
        pyr = B * C * P * e
        ref = B * C * hr * wr
        out_roi = K * R * C * self.PP * e
        n_gate = K * C * self.PP * e
        return {
            "fpn_gather_fwd": pyr + ref * e + 2 * ref,                       # + uint8 argmax (2 levels)
            "fpn_apply_fwd": 2 * pyr + ref * e + 2 * B * P * e,
            "roi_fuse_fwd": out_roi + pyr + 20 * K,
            "rff_gate_fwd": 4 * n_gate,
            "rff_gate_bwd": 6 * n_gate,
            "roi_fuse_bwd": out_roi + B * C * P * 4 + 20 * K,                # fp32 accumulators
            "fpn_apply_bwd": pyr + ref * e + 2 * B * P * e + ref * 4 + 2 * B * P * 4,
            "fpn_gather_bwd": ref * e + 2 * ref + pyr,
        }
