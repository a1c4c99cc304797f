#!/usr/bin/env python
"""Measure the other BASELINE.json configurations (the bench line itself is
configs[1]); prints one JSON object per case.  CUDA-event timing, pre-allocated
buffers, direct C-ABI calls (arfe_b200/workload.TrainStep).

  config 0 : Faster R-CNN inference, 1 image, K=1000 (fwd only)
  config 3 : RetinaNet neck only, B=8, strides 8..128 (AR-FPN fwd only)
  config 4 : Mask R-CNN bf16: bbox 7x7 x3 regions + mask 14x14 x1 region, fwd+bwd
  config 5 : Cascade: RoI fusion fwd+bwd per stage, K = 512..8192 per image
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arfe_b200 import _lib as L            # noqa: E402
from arfe_b200 import workload as wl       # noqa: E402

PEAK = 6542.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_ops(step, names, iters=20, warmup=5):
    """Average device ms of each named op of `step` (run in order)."""
    fns = {n: getattr(step, n) for n in names}
    ev = {n: [] for n in names}
    for it in range(warmup + iters):
        for n in names:
            if n == "roi_fuse_bwd":
                step.glue_before_roi_bwd()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            L.check(fns[n](), n)
            b.record()
            if it >= warmup:
                ev[n].append((a, b))
    torch.cuda.synchronize()
    return {n: sum(a.elapsed_time(b) for a, b in ev[n]) / len(ev[n]) for n in names}


def report(tag, step, ms, images):
    alg = step.algorithmic_bytes()
    total = sum(ms.values())
    out = {"case": tag, "images": images, "us_per_img": round(total * 1e3 / images, 1),
           "images_per_s": round(images / (total * 1e-3), 1),
           "ops": {n: {"ms": round(t, 4), "GBps": round(alg[n] / (t * 1e-3) / 1e9, 0),
                       "frac_of_measured_hbm": round(alg[n] / (t * 1e-3) / 1e9 / PEAK, 3)} for n, t in ms.items()}}
    print(json.dumps(out), flush=True)


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    FWD = ["fpn_gather_fwd", "fpn_apply_fwd", "roi_fuse_fwd", "rff_gate_fwd"]
    ROI = ["roi_fuse_fwd", "rff_gate_fwd", "rff_gate_bwd", "roi_fuse_bwd"]

    # config 0: inference, one image, K = 1000
    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        host = wl.host_inputs(1, 1000, 256, dtype=dt, channels_last=True)
        st = wl.TrainStep(host, dev)
        report(f"config0 inference 1 img K=1000 {name} (fwd)", st, time_ops(st, FWD), 1)
        del st, host

    # config 3: RetinaNet neck only, B = 8, strides 8..128
    host = wl.host_inputs(8, 16, 256, channels_last=True, strides=(8, 16, 32, 64, 128))
    st = wl.TrainStep(host, dev)
    report("config3 RetinaNet neck B=8 f32 (AR-FPN fwd)", st, time_ops(st, FWD[:2]), 8)
    report("config3 RetinaNet neck B=8 f32 (AR-FPN fwd+bwd)", st,
           time_ops(st, ["fpn_gather_fwd", "fpn_apply_fwd", "fpn_apply_bwd", "fpn_gather_bwd"]), 8)
    del st, host

    # config 4: Mask R-CNN bf16, bbox 7x7 x 3 regions and mask 14x14 x 1 region, 2 img x 512
    host = wl.host_inputs(2, 512, 256, dtype=torch.bfloat16, channels_last=True)
    st = wl.TrainStep(host, dev)
    report("config4 bf16 full step 2 img x 512 RoIs (7x7 x3)", st, time_ops(st, list(wl.KERNELS)), 2)
    del st, host
    host = wl.host_inputs(2, 128, 256, dtype=torch.bfloat16, channels_last=True, out_size=14)
    st = wl.TrainStep(host, dev, regions=1)
    report("config4 bf16 mask branch 2 img x 128 RoIs (14x14 x1) RoI part", st, time_ops(st, ROI[:1] + ROI[3:]), 2)
    del st, host

    # config 5: cascade stage, K sweep per image (one image per GPU, fwd + bwd of the RoI part)
    for K in (512, 1024, 2048, 4096, 8192):
        host = wl.host_inputs(1, K, 256, channels_last=True)
        st = wl.TrainStep(host, dev)
        report(f"config5 cascade stage K={K}/img f32 RoI part fwd+bwd", st, time_ops(st, ROI), 1)
        del st, host
    for K in (1024, 8192):
        host = wl.host_inputs(1, K, 256, dtype=torch.bfloat16, channels_last=True)
        st = wl.TrainStep(host, dev)
        report(f"config5 cascade stage K={K}/img bf16 RoI part fwd+bwd", st, time_ops(st, ROI), 1)
        del st, host
    # the reference layout through the compatibility kernels, for comparison
    host = wl.host_inputs(2, 512, 256)
    st = wl.TrainStep(host, dev, channels_last=False)
    report("configs[1] NCHW compatibility kernels f32 full step", st, time_ops(st, list(wl.KERNELS)), 2)


if __name__ == "__main__":
    main()
