#!/usr/bin/env python
"""bench.py -- headline benchmark of the region-aware feature path.

Workload (BASELINE.json configs[1], per GPU): one TRAINING step (fwd + bwd) of
fused AR-FPN + AR-RFF for 2 synthetic 800x1344 images x 512 RoIs (image-major,
as bbox2roi builds them), C=256, fp32, pyramid strides 4..64, RoI extractor on
strides 4..32.  metric = images/s (whole job, all GPUs); us_per_img is also
printed.  See arfe_b200/workload.py for the exact step.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA path, configs[1]
  python bench.py --config {0,2,3,4} [--rois-per-img K]       the other BASELINE configs
  python bench.py --impl reference ...                        reference CPU path

value : inputs resident in HBM, direct C-ABI calls, CUDA-event timed.
e2e   : same step through the C ABI, inputs start in PINNED HOST memory every
        step (H2D inside the timed region) and a result scalar is read back.
e2e_modules : the same step through the public modules / autograd Functions
        (what a swapped-in config runs), host buffers in, scalar out.
roofline : dominant kernel, algorithmic bytes / CUDA-event time vs the measured
        HBM copy peak (MEASURED_PEAKS.json).
verified : after the timed loop the buffers the timed steps wrote are compared
        with the oracle (the reference's arithmetic on the host): AR-FPN on all
        channels, the RoI part on all RoIs for a subset of channels (RoIAlign
        and the gate are independent per channel).
cpu_baseline : the oracle's CPU path (test infrastructure, used here only as
        the thing measured against) on a bounded sample, rank 0, N=1.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "images_per_sec", "images/s"
BATCH, ROIS_PER_IMG, CHANNELS = 2, 512, 256
WORKLOAD = ("configs[1]: Faster R-CNN R50 + AR-FPN + AR-RFF training step "
            "(fwd+bwd of the fused path), 2 img/GPU x 512 RoIs, 800x1344, C=256, "
            "strides 4-64, 3 regions, 7x7")
TRAFFIC_FILE = os.path.join("profiles", "r2_traffic.json")
VERIFY_CHANNELS = (0, 37, 101, 127, 128, 200, 254, 255)


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture; None when no
    capture exists for it."""
    for f in (TRAFFIC_FILE, os.path.join("profiles", "r1_traffic.json")):
        try:
            t = json.load(open(os.path.join(ROOT, f)))
            return float(t[kernel]["bytes"]), f
        except Exception:
            continue
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_rank_to_cores(local_rank, world):
    """Give each rank its own slice of the host cores the GPU is local to (NVML's ideal
    affinity), before any pinned buffer is allocated: staging memory is then first
    touched, and the copy threads run, next to the GPU's root port."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0))) or sorted(os.sched_getaffinity(0))
        per = max(len(allowed) // max(world, 1), 1)
        mine = allowed[(local_rank * per) % len(allowed):][:per] or allowed
        os.sched_setaffinity(0, mine)
        return {"gpu_ideal_cores": f"{cores[0]}-{cores[-1]}" if cores else None, "bound_to": f"{mine[0]}-{mine[-1]}"}
    except Exception as ex:  # pragma: no cover
        return {"error": str(ex)[:80]}


# ----------------------------------------------------------------- CPU arms
def cpu_sample_inputs(rois_per_img):
    from arfe_b200 import workload as wl
    return wl.host_inputs(batch=1, rois_per_img=rois_per_img, channels=CHANNELS, seed=0)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path
    (oracle/_ref = its RoIAlign sources compiled unmodified + the torch CPU ops
    its Python modules call), all host threads torch will use."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun pins OMP_NUM_THREADS=1; use every host core
    from oracle import arfe_oracle as O
    from oracle import build_oracle
    build_oracle.build_c_oracle()
    backend = "ref" if O.ref_ext() is not None else "c"
    kind = "reference" if backend == "ref" else "port"
    rpi = 64  # bounded sample: 1 image, 64 of its 512 RoIs (RoI part is linear in K)
    host = cpu_sample_inputs(rpi)
    for _ in range(args.warmup):
        O.reference_step(host, backend)
    t_fpn = t_roi = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = O.reference_step(host, backend)
        t_fpn += r["t_fpn"]
        t_roi += r["t_roi"]
    dt = (time.perf_counter() - t0) / args.steps
    t_img = t_fpn / args.steps + t_roi / args.steps * (ROIS_PER_IMG / rpi)
    value = 1.0 / t_img
    cores = torch.get_num_threads()
    sample = (f"per step: 1 image 800x1344 C=256 with {rpi} of its {ROIS_PER_IMG} RoIs; value = 1/(t_fpn + "
              f"{ROIS_PER_IMG // rpi} x t_roi) (the RoI part is a per-RoI loop, linear in K); the reference's "
              f"RoIAlign loop is single-threaded (roi_align_v2.cpp:115-117), torch ops use {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "us_per_img": 1e6 / value, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "timing": "host wall clock, CPU only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Bounded CPU sample for our arm's JSON line (rank 0, N=1)."""
    try:
        from oracle import arfe_oracle as O
        from oracle import build_oracle
        build_oracle.build_c_oracle()
    except Exception as ex:  # pragma: no cover
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    torch.set_num_threads(os.cpu_count() or 1)
    backend, kind = "c", "port"
    rpi = 128
    host = cpu_sample_inputs(rpi)
    O.reference_step(host, backend)  # warm
    t0 = time.perf_counter()
    n, t_fpn, t_roi = 0, 0.0, 0.0
    while n < 2 or time.perf_counter() - t0 < 10.0:
        r = O.reference_step(host, backend)
        t_fpn, t_roi, n = t_fpn + r["t_fpn"], t_roi + r["t_roi"], n + 1
    t_img = t_fpn / n + t_roi / n * (ROIS_PER_IMG / rpi)
    cores = os.cpu_count() or 1
    return {"value": 1.0 / t_img, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": (f"{n} steps of 1 image 800x1344 C=256 with {rpi} of its {ROIS_PER_IMG} RoIs; value = "
                       f"1/(t_fpn + {ROIS_PER_IMG // rpi} x t_roi); oracle C port (OpenMP forward over {cores} "
                       f"cores, serial backward) + torch CPU ops on {torch.get_num_threads()} threads")}


# ------------------------------------------------------------- verification
def verify_step(step, host):
    """Compare what the timed steps left in the device buffers with the oracle
    (test infrastructure used as the checker).  Returns {"verified": bool, ...}."""
    from oracle import arfe_oracle as O
    from oracle import build_oracle
    build_oracle.build_c_oracle()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    S = [c for c in VERIFY_CHANNELS if c < step.C]
    cpu = lambda t: t.detach().float().cpu().contiguous()
    nchw = lambda v: [nchw(t) for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else \
        (v.contiguous() if torch.is_tensor(v) else v)
    dy_dev = [cpu(t) for t in step.dy]
    ref = O.reference_step({k: nchw(v) for k, v in host.items()}, backend="c", channels=S,
                           dy_full=dy_dev, roi_levels=step.rlev)
    worst, failed = {}, []

    def check(name, got, want, rel, scale):
        got, want = cpu(got), want.detach().float()
        err = (got - want).abs()
        tol = rel * want.abs() + scale * float(want.abs().max()) + 1e-6
        ratio = float((err / tol).max()) if err.numel() else 0.0
        worst[name] = max(worst.get(name, 0.0), round(ratio, 3))
        if ratio > 1.0 or not bool(torch.isfinite(got).all()):
            failed.append(name)

    exact = bool(torch.equal(cpu(step.gathered), ref["gathered"]))
    if not exact:
        failed.append("gathered (bit-exact)")
    F = [cpu(t) for t in step.Fr] if step.split else \
        [cpu(step.F[:, r * step.C:(r + 1) * step.C]) for r in range(step.R)]
    c = len(S)
    d_ori = step.d_ori if step.split else step.dF[:, :step.C]
    for l in range(step.nlev):
        check("y", step.y[l], ref["y"][l], 1e-5, 0.0)
        check("dy", dy_dev[l][:, S], ref["dy"][l], 1e-5, 2e-5)
        check("dx", cpu(step.dx[l]), ref["dx"][l], 1e-5, 2e-5)
        check("dg1", step.dg1[l], ref["dg1"][l], 1e-5, 2e-5)
        check("dg2", step.dg2[l], ref["dg2"][l], 1e-5, 2e-5)
    for r in range(step.R):
        check("F", F[r][:, S], ref["F"][:, r * c:(r + 1) * c], 1e-5, 0.0)
    check("z", cpu(step.z)[:, S], ref["z"], 1e-5, 0.0)
    check("d_ori", cpu(d_ori)[:, S], ref["d_ori"], 1e-5, 0.0)
    check("d_ab", cpu(step.d_ab)[:, S], ref["d_ab"], 1e-5, 0.0)
    check("dbsf", step.dbsf, ref["dbsf"], 1e-5, 2e-5)
    return {"verified": not failed, "failed": failed, "gather_bit_exact": exact,
            "max_err_over_tol": worst, "channels_checked_roi_part": S,
            "tolerance": "|err| <= 1e-5 |ref| + 1e-6 (forward); + 2e-5 max|ref| for gradients summed in another order",
            "against": "oracle.reference_step (C port of roi_align_v2.cpp + torch CPU ops) on the same host inputs, "
                       "all RoIs, all pixels; d x / d bsf / d gate maps from the device d y",
            "seconds": round(time.perf_counter() - t0, 1)}


# ------------------------------------------------------------------ GPU arm
def make_timer(names, only=None):
    from arfe_b200 import _lib as L
    ev = {n: [] for n in names}

    def timed(name, fn):
        if only is not None and name != only:
            L.check(fn(), name)
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.check(fn(), name)
        b.record()
        ev[name].append((a, b))
    return ev, timed


def mean_ms(ev):
    return {n: sum(a.elapsed_time(b) for a, b in v) / len(v) for n, v in ev.items() if v}


def time_case(case, steps, warmup, barrier=None):
    """Device-timed steps of a schedule: (ms per step, per-op ms, dominant op, its ms inside the run)."""
    names = case.op_names()
    for _ in range(max(warmup - 2, 1)):
        case.step()
    ev_w, timed_w = make_timer(names)
    for _ in range(2):
        case.step(timed_w)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    top = max(mean_ms(ev_w).items(), key=lambda kv: kv[1])[0]
    ev_top, timed_top = make_timer(names, only=top)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        case.step(timed_top)
    b.record()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    ms_total = a.elapsed_time(b)
    top_ms = mean_ms(ev_top)[top]
    ev_all, timed_all = make_timer(names)
    for _ in range(steps):
        case.step(timed_all)
    torch.cuda.synchronize()
    return ms_total, mean_ms(ev_all), top, top_ms


def kernel_table(alg, per_ms, peak):
    return {n: {"ms": round(per_ms[n], 4), "alg_MB": round(alg[n] / 1e6, 1),
                "GBps": round(alg[n] / (per_ms[n] * 1e-3) / 1e9, 1),
                "frac": round(alg[n] / (per_ms[n] * 1e-3) / 1e9 / peak, 3)} for n in per_ms}


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops"]), "measured burst cuBLAS bf16 (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md)"


def nonlocal_refine_block(dev, verify=True):
    """SURVEY 8(f) row 1, next to the path: the NonLocal2D attention of the refine level of configs[1]
    (2 images, 256 channels, 50 x 84 positions) -- fused tensor-core kernel (pack + attention + merge,
    the whole arfe_nonlocal_attention_forward call) against the reference's two matmuls + softmax through
    the library in fp32, both device-timed on resident inputs; checked against the oracle."""
    import arfe_b200 as A
    B, D, H, W = 2, 256, 50, 84
    g = torch.Generator().manual_seed(5)
    host = [torch.randn(B, D, H, W, generator=g) * s for s in (0.25, 0.25, 1.0)]
    ts = [t.to(dev).contiguous(memory_format=torch.channels_last) for t in host]

    def lib_ref():
        th = ts[0].reshape(B, D, -1).permute(0, 2, 1)
        pw = torch.matmul(th, ts[1].reshape(B, D, -1)).softmax(dim=-1)
        return torch.matmul(pw, ts[2].reshape(B, D, -1).permute(0, 2, 1))

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_f = timeit(lambda: A.nonlocal_attention(*ts))
    tb = [t.bfloat16() for t in ts]          # bf16 channels-last: read in place through tensor maps, no packing pass
    t_b = timeit(lambda: A.nonlocal_attention(*tb))
    t_l = timeit(lib_ref, 5)
    flops = 4.0 * B * (H * W) ** 2 * D
    peak, src = tensor_peak()
    out = {"shape": f"B={B} C={D} {H}x{W} ({H * W} positions), fp32 in/out, operands bf16, channels-last",
           "fused_us": round(t_f * 1e3, 1), "fused_bf16_io_us": round(t_b * 1e3, 1), "library_fp32_us": round(t_l * 1e3, 1),
           "tflops": round(flops / (t_f * 1e-3) / 1e12, 1), "peak_tflops": peak, "peak_source": src,
           "frac": round(flops / (t_f * 1e-3) / 1e12 / peak, 3),
           "frac_bf16_io": round(flops / (t_b * 1e-3) / 1e12 / peak, 3), "gpu_launches_per_call": 2,
           "timing": "CUDA events around 20 calls through the Python wrapper (the call replayed as a CUDA graph is 3-4 us shorter)"}
    if verify:
        from oracle import arfe_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        want = O.nonlocal_attention(*host)
        got = A.nonlocal_attention(*ts).float().cpu()
        err = float((got - want).abs().max() / want.abs().max())
        out["verified"] = bool(err <= 1e-2 and torch.isfinite(got).all())
        out["max_err_over_max_ref"] = round(err, 5)
        out["tolerance"] = "|err| <= 1e-2 max|ref| (bf16 operands, north_star's bf16 bound) against oracle.nonlocal_attention"
    return out


def neck_module_block(dev):
    """Module level: WFPNDualSpatial (the reference's neck class) on the configs[1] pyramid, channels-last fp32,
    forward and forward + backward, with the refine block's attention through the library and through the
    fused tensor-core kernel -- what swapping the module into a detector buys, device-timed, inputs resident."""
    import arfe_b200 as A
    from arfe_b200 import workload as wl
    shapes = wl.pyramid_shapes(800, 1344)
    B, C = 2, 256
    out = {"shape": f"B={B} C={C} 5 levels of 800x1344, fp32 channels-last; gather / gate convolutions / gated residual fused in both arms"}

    def timeit(fn, n):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    g = torch.Generator().manual_seed(6)
    xs = [torch.randn(B, C, h, w, generator=g).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
          for h, w in shapes]
    gs = [torch.randn_like(x) for x in xs]
    m = A.WFPNDualSpatial(C, 5).to(dev).to(memory_format=torch.channels_last)
    m.init_weights()
    torch.nn.init.normal_(m.refine.conv_out.conv.weight, 0, 0.02)
    for tag, fused in (("library_attention", False), ("fused_attention", True)):
        m.refine.fused_attention = fused

        def fwd():
            with torch.no_grad():
                return m(xs)

        def fwdbwd():
            torch.autograd.backward(list(m(xs)), gs)

        out[tag] = {"forward_ms": round(timeit(fwd, 10), 3), "forward_backward_ms": round(timeit(fwdbwd, 5), 3)}
    return out


def other_configs_block(dev, peak):
    """Compact device-timed numbers of the other BASELINE configs (the full lines come
    from --config N): tracked by the driver round over round."""
    from arfe_b200 import workload as wl
    out = {}
    for tag, cfg, kpi in (("config0_infer_K1000", 0, None), ("config2_retina_neck_B8", 2, None),
                          ("config3_mask_bf16", 3, None), ("config4_cascade_K512", 4, 512),
                          ("config4_cascade_K8192", 4, 8192)):
        try:
            case = wl.make_case(cfg, dev, rois_per_img=kpi)
            ms_total, per_ms, top, _ = time_case(case, 10, 3)
            alg = case.algorithmic_bytes()
            ms = ms_total / 10
            out[tag] = {"us_per_img": round(ms * 1e3 / case.images, 1),
                        "images_per_s": round(case.images / (ms * 1e-3), 1), "dtype": case.dtype,
                        "step_frac_of_hbm_peak": round(sum(alg.values()) / (ms * 1e-3) / 1e9 / peak, 3),
                        "slowest_op": top, "slowest_op_ms": round(per_ms[top], 4),
                        "slowest_op_frac": round(alg[top] / (per_ms[top] * 1e-3) / 1e9 / peak, 3)}
            del case
            torch.cuda.empty_cache()
        except Exception as ex:  # pragma: no cover
            out[tag] = {"error": str(ex)[:200]}
    return out


def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from arfe_b200 import _lib as L
    from arfe_b200 import workload as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    affinity = bind_rank_to_cores(local_rank, world)
    L.lib()  # fail loudly if the extension is missing
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peak, peak_src = peaks()
    cl = not args.nchw
    main_cfg = args.config == 1
    if main_cfg:
        host = wl.host_inputs(BATCH, args.rois_per_img or ROIS_PER_IMG, CHANNELS, seed=rank, pin=True,
                              channels_last=cl)
        step = wl.TrainStep(host, dev, channels_last=cl)
        case = wl.Case(WORKLOAD, BATCH, [("", step, step.kernels)], "f32")
    else:
        case = wl.make_case(args.config, dev, seed=rank, rois_per_img=args.rois_per_img)
        step = case.parts[0][1]
    images = case.images

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, per_kernel_ms, top, top_ms = time_case(case, args.steps, args.warmup, barrier)
    clocks = sampler.stop() if rank == 0 else None

    verification = None
    if main_cfg and rank == 0 and not args.no_verify:
        # the buffers still hold what the last timed step wrote
        verification = verify_step(step, host)

    e2e = e2e_mod = None
    if main_cfg:
        e2e_ms, h2d, d2h, h2d_gbps = run_e2e(args, host, dev, barrier, cl)
        mod_ms, mod_dev_ms = run_e2e_modules(args, host, dev, barrier) if cl else (None, None)
        t = torch.tensor([ms_total, e2e_ms, mod_ms or 0.0, -h2d_gbps[0], -h2d_gbps[1], mod_dev_ms or 0.0],
                         dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, mod_ms = float(t[0]), float(t[1]), (float(t[2]) if mod_ms else None)
        mod_dev_ms = float(t[5]) if mod_dev_ms else None
        e2e = {"value": world * images / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
               "h2d_GBps_per_gpu": {"one_rank_at_a_time": round(-float(t[3]), 1),
                                    "all_ranks_concurrently": round(-float(t[4]), 1),
                                    "note": "slowest rank; pinned host -> device copy of one step's inputs, outside the timed region"},
               "host_affinity": affinity,
               "note": "pinned-host pyramid + RoIs + conv stand-ins copied H2D every step (double-buffered: the copies of step "
                       "k+1 overlap step k), loss scalar read back every step; bound by the H2D copy"}
        if mod_ms:
            e2e_mod = {"value": world * images / (mod_ms / args.steps * 1e-3), "unit": UNIT,
                       "ms_per_step": mod_ms / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "device_resident": {"value": world * images / (mod_dev_ms / args.steps * 1e-3), "unit": UNIT,
                                           "ms_per_step": mod_dev_ms / args.steps,
                                           "note": "same calls, inputs already in HBM, no read-back: compare with `value` "
                                                   "(pre-planned, pre-allocated C-ABI calls) for the cost of the module path"},
                       "path": "arfe_b200.fpn_gather / fpn_apply / roi_fuse_split / rff_gate autograd Functions "
                               "(what WFPNDualSpatial.forward and StandardRoIHead._bbox_forward call; the convs "
                               "between them are stand-in inputs as in `value`), torch.autograd.backward, "
                               "per-call allocation, tile bins on a side stream"}
    else:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * images / (ms_step * 1e-3)
        alg = case.algorithmic_bytes()
        achieved = alg[top] / (top_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(top) if (cl and main_cfg) else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "us_per_img": 1e6 / value * world,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": case.dtype,
            "data": "synthetic",
            "config": {"workload": case.workload, "baseline_config": args.config,
                       "global_batch": world * images, "rois_per_gpu": step.K,
                       "roi_order": "image-major (per-image blocks, bbox2roi)",
                       "roi_levels": f"extractor on the first {step.rlev} of {step.nlev} pyramid levels (featmap_strides 4-32)",
                       "parallelism": f"dp{world} (images sharded, no data-path collective)",
                       "memory_format": "torch.channels_last (fast path)" if cl else "NCHW (reference layout, compatibility kernels)",
                       "roi_tensors": ("regions as separate tensors (ori | lw | lh), read/written in place: no cat / slice copies in the step"
                                       if step.split else "concatenated [K, 3C, 7, 7]; torch slice copies between the kernels are inside the step"),
                       "backward": ("AR-FPN backward fused (arfe_fpn_backward_fused: dbsf, d gate maps and d x_l = d out_l + gather "
                                    "gradient, d out read once); " if cl else
                                    "d x_l = d out_l + gather gradient in one write (arfe_fpn_gather_backward_acc); ") +
                                   "d lw / d lh are separate input tensors",
                       "streams": ("the RoI plan (under the AR-FPN forward kernels) and the backward's tile binning (forked after the "
                                   "RoI forward, under the gate kernels) run on a second stream; the per-op times of roi_fuse_fwd / "
                                   "roi_fuse_bwd exclude them, the ops they overlap and the step time include them"
                                   if getattr(step, "overlap_plan", False) else "one stream"),
                       "l2": "inputs+outputs per step (~1.8 GB) exceed the 126 MB L2; no explicit flush",
                       "timing": "CUDA events on the launch stream, max over ranks",
                       "instrumentation": "the timed region carries events around the dominant kernel only (roofline.achieved); "
                                          "`kernels` is a second pass of the same K steps with events around every op"},
            "clocks": clocks,
            "gpu_launches": case.launches_per_step() * args.steps,
            "roofline": {"bound": "hbm", "kernel": top, "kernel_ms": top_ms, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": alg[top],
                         "peak_source": peak_src,
                         "traffic_source": f"ncu --set full capture, {traffic_src}" if traffic_src else None,
                         "step_algorithmic_bytes": sum(alg.values()),
                         "step_frac": sum(alg.values()) / (ms_step * 1e-3) / 1e9 / peak},
            "kernels": kernel_table(alg, per_kernel_ms, peak),
        }
        if e2e:
            line["e2e"] = e2e
        if e2e_mod:
            line["e2e_modules"] = e2e_mod
        if verification is not None:
            line["verified"] = verification["verified"]
            line["verification"] = verification
        if main_cfg and world == 1 and not args.no_other_configs:
            line["other_configs"] = other_configs_block(dev, peak)
            try:
                line["next_rows"] = {"nonlocal_refine": nonlocal_refine_block(dev, verify=not args.no_verify),
                                     "neck_module": neck_module_block(dev)}
            except Exception as ex:  # pragma: no cover
                line["next_rows"] = {"nonlocal_refine": {"error": str(ex)[:200]}}
        if main_cfg and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


E2E_KEYS_LIST = ("x", "g1", "g2")
E2E_KEYS = ("bsf", "rois", "a", "b", "gz", "glw", "glh", "gbsf")


def run_e2e(args, host, dev, barrier, cl):
    """Same step, but every step starts from pinned host memory (H2D inside the
    timed region, every step) and ends with a device->host read of the result
    scalar.  The copies of step k+1 run on a second stream into the other of two
    input buffer sets while step k computes (what any input pipeline does); the
    region is bound by the copy either way."""
    from arfe_b200 import workload as wl
    steps = [wl.TrainStep(host, dev, channels_last=cl) for _ in range(2)]

    def pairs_of(step):
        pairs = []
        for key in E2E_KEYS_LIST:
            pairs += list(zip(getattr(step, key), host[key]))
        for key in E2E_KEYS:
            pairs.append((getattr(step, key), host[key]))
        return pairs
    pairs = [pairs_of(s) for s in steps]
    h2d = sum(s.numel() * s.element_size() for _, s in pairs[0])
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream(dev)
    copier = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]   # inputs of buffer set i have landed
    done = [torch.cuda.Event() for _ in range(2)]    # the step on buffer set i has finished
    used = [False, False]

    def issue_copy(k):
        i = k % 2
        with torch.cuda.stream(copier):
            if used[i]:
                copier.wait_event(done[i])
            for d, s in pairs[i]:
                d.copy_(s, non_blocking=True)
            ready[i].record(copier)

    def run(n):
        issue_copy(0)
        for k in range(n):
            i = k % 2
            if k + 1 < n:
                issue_copy(k + 1)
            main.wait_event(ready[i])
            steps[i].step()
            loss_host.copy_(steps[i].z.sum() + steps[i].dx[4].sum(), non_blocking=True)
            done[i].record(main)
            used[i] = True
            main.synchronize()   # the result scalar is on the host

    # what the copy alone achieves: this rank by itself, then all ranks at once
    def h2d_rate(n=3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            for d, s in pairs[0]:
                d.copy_(s, non_blocking=True)
        b.record()
        torch.cuda.synchronize(dev)
        return n * h2d / (a.elapsed_time(b) * 1e-3) / 1e9
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    h2d_rate(1)
    alone = 0.0
    for r in range(world):
        barrier()
        if r == rank:
            alone = h2d_rate()
    barrier()
    together = h2d_rate()

    run(min(args.warmup, 3))
    torch.cuda.synchronize(dev)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(args.steps)
    b.record()
    barrier()
    return a.elapsed_time(b), h2d, 4, (alone, together)


def run_e2e_modules(args, host, dev, barrier):
    """The step through the public autograd Functions (the calls the modules make), host
    buffers in, scalar out, every step: per-call allocation, autograd bookkeeping, the
    gradient accumulation of x (residual + gather paths) done by autograd."""
    import arfe_b200 as A
    strides = host["strides"]
    scales = [1.0 / s for s in strides[:4]]
    cl = lambda t: t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t
    keys = E2E_KEYS_LIST + E2E_KEYS
    bufs = [{k: ([torch.empty_like(cl(t), device=dev) for t in host[k]] if isinstance(host[k], list)
                 else torch.empty_like(cl(host[k]), device=dev)) for k in keys} for _ in range(2)]
    main = torch.cuda.current_stream(dev)
    copier = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    used = [False, False]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def issue_copy(k):
        i = k % 2
        with torch.cuda.stream(copier):
            if used[i]:
                copier.wait_event(done[i])
            for key in keys:
                if isinstance(host[key], list):
                    for d, s in zip(bufs[i][key], host[key]):
                        d.copy_(s, non_blocking=True)
                else:
                    bufs[i][key].copy_(host[key], non_blocking=True)
            ready[i].record(copier)

    def one(b):
        x = [t.detach().requires_grad_(True) for t in b["x"]]
        bsf = b["bsf"].detach().requires_grad_(True)
        g1 = [t.detach().requires_grad_(True) for t in b["g1"]]
        g2 = [t.detach().requires_grad_(True) for t in b["g2"]]
        a_, b_ = b["a"].detach().requires_grad_(True), b["b"].detach().requires_grad_(True)
        gathered = A.fpn_gather(x, 2)
        y = A.fpn_apply(x, bsf, g1, g2)
        ori, lw, lh = A.roi_fuse_split(list(y[:4]), b["rois"], 7, scales, regions=3)
        z = A.rff_gate(ori, a_, b_)
        torch.autograd.backward([z, lw, lh, gathered], [b["gz"], b["glw"], b["glh"], b["gbsf"]])
        return z.detach().sum() + x[4].grad.sum()

    def run(n):
        issue_copy(0)
        for k in range(n):
            i = k % 2
            if k + 1 < n:
                issue_copy(k + 1)
            main.wait_event(ready[i])
            loss_host.copy_(one(bufs[i]), non_blocking=True)
            done[i].record(main)
            used[i] = True
            main.synchronize()

    run(min(args.warmup, 3))
    torch.cuda.synchronize(dev)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(args.steps)
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    # the same calls with the inputs already resident in HBM (no copies, no read-back)
    for _ in range(3):
        one(bufs[0])
    barrier()
    a.record()
    for _ in range(args.steps):
        one(bufs[0])
    b.record()
    barrier()
    return e2e_ms, a.elapsed_time(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4],
                    help="index into BASELINE.json configs (default 1: the headline training step)")
    ap.add_argument("--rois-per-img", type=int, default=None, help="RoIs per image (config 4 sweep: 512 ... 8192)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="reference memory layout through the compatibility kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
