"""RoI forward / backward of the bench workload (configs[1] shapes): CUDA-event times
of the calls with a ready plan (what bench.py's per-op table shows) and with plan / bins
inside; also the ncu target (`python scripts/roi_probe.py ncu` runs 3 plain iterations).
Optional env: ARFE_B200_LIB=arfe_b200/libarfe_b200_prof.so + the knob variables."""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
K = int(os.environ.get("PROBE_K", "512"))
B = int(os.environ.get("PROBE_B", "2"))
dt = torch.bfloat16 if os.environ.get("PROBE_BF16") else torch.float32
host = wl.host_inputs(B, K, 256, channels_last=True, device=dev, dtype=dt)
if os.environ.get("PROBE_SORT"):
    # RoIs ordered by (image, level of the original box, top edge): a locality probe
    r = host["rois"]
    s_ = torch.sqrt((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2]))
    lv = torch.floor(torch.log2(s_ / 56 + 1e-6)).clamp(0, 3)
    key = r[:, 0] * 1e6 + (3 - lv) * 1e5 + r[:, 2] * (1.0 / torch.pow(2.0, lv)) 
    host["rois"] = r[torch.argsort(key)].contiguous()
st = wl.TrainStep(host, dev)
st.step()
torch.cuda.synchronize()


def fwd(ready):
    st.async_plan = 0
    if ready:
        st.plan_async(bins=False)
    L.check(st.roi_fuse_fwd(), "fwd")


def bwd(ready):
    st.async_bins = 0
    st.planned = 1
    if ready:
        st.plan_async(bins=True)
        st.bins_async()
        st.async_plan = 0
    L.check(st.roi_fuse_bwd(), "bwd")


def timeit(fn, ready, n=30):
    for _ in range(5):
        fn(ready)
    torch.cuda.synchronize()
    ev = []
    for _ in range(n):
        if ready:  # the plan / bins are built outside the bracket
            fn_a, fn_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if fn is fwd:
                st.async_plan = 0
                st.plan_async(bins=False)
                torch.cuda.current_stream().wait_event(st.ev_plan)
                st.async_plan = 0
                fn_a.record()
                L.check(st.lib.arfe_roi_fuse_forward_plan_split(
                    st.p_y, st.H, st.W, st.scales, st.rlev, st.B, st.C, st.rois.data_ptr(), st.K, st.R, 1.0,
                    st.P, st.P, 0, 56.0, st.dt, st.p_Fr, st.ws_ptr, st.ws_bytes, 1, st.stream), "fwd")
                fn_b.record()
            else:
                st.plan_async(bins=True)
                st.bins_async()
                torch.cuda.current_stream().wait_event(st.ev_bin)
                st.async_plan = st.async_bins = 0
                fn_a.record()
                L.check(st.lib.arfe_roi_fuse_backward_pull_split(
                    st.p_dFr, st.H, st.W, st.scales, st.rlev, st.B, st.C, st.rois.data_ptr(), st.K, st.R, 1.0,
                    st.P, st.P, 0, 56.0, st.dt, st.p_dy, st.ws_ptr, st.ws_bytes, 2, st.stream), "bwd")
                fn_b.record()
            ev.append((fn_a, fn_b))
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(False)
            b.record()
            ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3


def digest(ts):
    h = hashlib.sha256()
    for t in ts:
        h.update(t.float().cpu().numpy().tobytes())
    return h.hexdigest()[:12]


if mode == "ncu":
    for _ in range(3):
        fwd(False)
        bwd(False)
    torch.cuda.synchronize()
else:
    tag = os.environ.get("PROBE_TAG", "")
    print(f"{tag} K={K}/img B={B} {str(dt)[6:]}: fwd ready-plan {timeit(fwd, True):7.1f} us  fwd+plan {timeit(fwd, False):7.1f} us | "
          f"bwd ready-bins {timeit(bwd, True):7.1f} us  bwd+bin {timeit(bwd, False):7.1f} us | "
          f"F {digest(st.Fr)} dy {digest(st.dy[:st.rlev])}", flush=True)
