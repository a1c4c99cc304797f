#!/usr/bin/env python
"""bench.py -- headline benchmark of the region-aware feature path.

Workload (BASELINE.json configs[1], per GPU): one TRAINING step (fwd + bwd) of
fused AR-FPN + AR-RFF for 2 synthetic 800x1344 images x 512 RoIs, C=256, fp32,
FPN strides 4..64.  metric = images/s (whole job, all GPUs); us_per_img is
also printed.  See arfe_b200/workload.py for the exact step.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA path
  python bench.py --impl reference ...                        reference CPU path

value : inputs resident in HBM, direct C-ABI calls, CUDA-event timed.
e2e   : same step, inputs start in PINNED HOST memory every step (H2D inside
        the timed region) and a result scalar is read back (D2H).
roofline : dominant kernel, algorithmic bytes / CUDA-event time vs the measured
        HBM copy peak (MEASURED_PEAKS.json).
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


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture
    (profiles/r1_traffic.json); None when no capture exists for it."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return float(t[kernel]["bytes"])
    except Exception:
        return None


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


# ----------------------------------------------------------------- CPU arms
def cpu_step(oracle, host, backend):
    """The same step on the host through the oracle (reference arithmetic).
    Returns (seconds in the AR-FPN part, seconds in the AR-RFF part)."""
    C = host["a"].shape[1]
    t0 = time.perf_counter()
    x = [t.clone().requires_grad_(True) for t in host["x"]]
    bsf = host["bsf"].clone().requires_grad_(True)
    g1 = [t.clone().requires_grad_(True) for t in host["g1"]]
    g2 = [t.clone().requires_grad_(True) for t in host["g2"]]
    gathered = oracle.wfpn_gather(x, 2)
    y = oracle.wfpn_apply(x, bsf, g1, g2)
    t1 = time.perf_counter()
    yd = [t.detach().requires_grad_(True) for t in y]
    F = oracle.arrff_bbox_feats(yd, host["rois"], [4, 8, 16, 32, 64], backend=backend)
    a = host["a"].clone().requires_grad_(True)
    b = host["b"].clone().requires_grad_(True)
    z = oracle.rff_gate(F[:, :C], a, b)
    # gradients reaching the two context-region blocks: synthetic stand-in for
    # the conv backward that stays on PyTorch (same role as TrainStep's glue)
    torch.autograd.backward([z, F[:, C:2 * C], F[:, 2 * C:]], [host["gz"], host["gz"], host["gz"]])
    t2 = time.perf_counter()
    dy = [t.grad if t.grad is not None else torch.zeros_like(t) for t in yd]
    torch.autograd.backward(list(y) + [gathered], dy + [host["gbsf"]])
    t3 = time.perf_counter()
    return (t1 - t0) + (t3 - t2), (t2 - t1)


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
        cpu_step(O, host, backend)
    t_fpn = t_roi = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f, r = cpu_step(O, host, backend)
        t_fpn += f
        t_roi += r
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
    cpu_step(O, host, backend)  # warm
    t0 = time.perf_counter()
    n, t_fpn, t_roi = 0, 0.0, 0.0
    while n < 2 or time.perf_counter() - t0 < 10.0:
        f, r = cpu_step(O, host, backend)
        t_fpn, t_roi, n = t_fpn + f, t_roi + r, n + 1
    t_img = t_fpn / n + t_roi / n * (ROIS_PER_IMG / rpi)
    cores = os.cpu_count() or 1
    return {"value": 1.0 / t_img, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": (f"{n} steps of 1 image 800x1344 C=256 with {rpi} of its {ROIS_PER_IMG} RoIs; value = "
                       f"1/(t_fpn + {ROIS_PER_IMG // rpi} x t_roi); oracle C port (OpenMP forward over {cores} "
                       f"cores, serial backward) + torch CPU ops on {torch.get_num_threads()} threads")}


# ------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from arfe_b200 import _lib as L
    from arfe_b200 import workload as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    L.lib()  # fail loudly if the extension is missing
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cl = not args.nchw
    host = wl.host_inputs(BATCH, ROIS_PER_IMG, CHANNELS, seed=rank, pin=True, channels_last=cl)
    step = wl.TrainStep(host, dev, channels_last=cl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident steps; CUDA events around the dominant kernel inside the timed
    # region.  Events around all eight ops cost 38 us per step (scripts/step_modes.py:
    # 629 us plain, 667 us instrumented), so the per-op table is a separate pass.
    names = wl.KERNELS

    def make_timer(only=None):
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

    for _ in range(max(args.warmup - 2, 1)):
        step.step()
    ev_w, timed_w = make_timer()      # last warm-up steps: which op is the dominant one
    for _ in range(2):
        step.step(timed_w)
    barrier()
    top = max(mean_ms(ev_w).items(), key=lambda kv: kv[1])[0]
    ev_top, timed_top = make_timer(only=top)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step.step(timed_top)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_start.elapsed_time(t_end)
    top_ms = mean_ms(ev_top)[top]
    # per-op table: the same K steps again with events around every op (outside the timed region)
    ev_all, timed_all = make_timer()
    for _ in range(args.steps):
        step.step(timed_all)
    barrier()
    per_kernel_ms = mean_ms(ev_all)

    # ---- e2e: host buffers in, scalar out, public module-level API --------
    e2e_ms, h2d, d2h = run_e2e(args, host, dev, barrier, cl)

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * BATCH / (ms_step * 1e-3)
        e2e_value = world * BATCH / (e2e_ms / args.steps * 1e-3)
        alg = step.algorithmic_bytes()
        peak, peak_src = peaks()
        achieved = alg[top] / (top_ms * 1e-3) / 1e9
        kern = {n: {"ms": round(per_kernel_ms[n], 4), "alg_MB": round(alg[n] / 1e6, 1),
                    "GBps": round(alg[n] / (per_kernel_ms[n] * 1e-3) / 1e9, 1),
                    "frac": round(alg[n] / (per_kernel_ms[n] * 1e-3) / 1e9 / peak, 3)} for n in names}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "us_per_img": 1e6 / value * world,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "rois_per_gpu": BATCH * ROIS_PER_IMG,
                       "parallelism": f"dp{world} (images sharded, no data-path collective)",
                       "memory_format": "torch.channels_last (fast path)" if cl else "NCHW (reference layout, compatibility kernels)",
                       "roi_tensors": ("regions as separate tensors (ori | lw | lh), read/written in place: no cat / slice copies in the step"
                                       if step.split else "concatenated [K, 3C, 7, 7]; torch slice copies between the kernels are inside the step"),
                       "streams": ("RoI plan and tile binning (2 kernels, ~40 us) run on a second stream under the AR-FPN forward kernels; "
                                   "the per-op times of roi_fuse_fwd / roi_fuse_bwd exclude them, the step time includes them"
                                   if getattr(step, "overlap_plan", False) else "one stream"),
                       "l2": "inputs+outputs per step (~1.5 GB) exceed the 126 MB L2; no explicit flush",
                       "timing": "CUDA events on the launch stream, max over ranks",
                       "instrumentation": "the timed region carries events around the dominant kernel only (roofline.achieved); "
                                          "`kernels` is a second pass of the same K steps with events around every op (+38 us per step)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "note": "pinned-host pyramid+RoIs+conv stand-ins copied H2D every step (double-buffered: the copies of step k+1 overlap step k), loss scalar read back every step; PCIe-bound"},
            "gpu_launches": step.launches_per_step() * args.steps,
            "roofline": {"bound": "hbm", "kernel": top, "kernel_ms": top_ms, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(top) if cl else None,
                         "algorithmic_bytes": alg[top], "peak_source": peak_src,
                         "traffic_source": "ncu --set full capture, profiles/r1_traffic.json"},
            "kernels": kern,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, host, dev, barrier, cl):
    """Same step, but every step starts from pinned host memory (H2D inside the
    timed region, every step) and ends with a device->host read of the result
    scalar.  The copies of step k+1 run on a second stream into the other of two
    input buffer sets while step k computes (what any input pipeline does); the
    region is PCIe-bound either way (356 MB per step)."""
    from arfe_b200 import workload as wl
    steps = [wl.TrainStep(host, dev, channels_last=cl) for _ in range(2)]

    def pairs_of(step):
        pairs = []
        for key in ("x", "g1", "g2"):
            pairs += list(zip(getattr(step, key), host[key]))
        for key in ("bsf", "rois", "a", "b", "gz", "gbsf"):
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

    run(min(args.warmup, 3))
    torch.cuda.synchronize(dev)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(args.steps)
    b.record()
    barrier()
    return a.elapsed_time(b), h2d, 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
