"""Time the RoI forward of the bench workload under the ARFE_FWD_SKIP profiling
knobs (1: consumers skip the math, 2: producer skips the copies, 16: static
striding instead of the cost-balanced partition)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L
dev = torch.device("cuda:0")
host = wl.host_inputs(2, 512, 256, channels_last=True)
st = wl.TrainStep(host, dev)
def run(knob, n=30):
    os.environ["ARFE_FWD_SKIP"] = str(knob)
    for _ in range(5): L.check(st.roi_fuse_fwd(), "f")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): L.check(st.roi_fuse_fwd(), "f")
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for knob in [int(x) for x in (sys.argv[1:] or ["0", "16", "1", "17", "2", "18", "3", "19"])]:
    print(f"ARFE_FWD_SKIP={knob:3d}: {run(knob):8.1f} us")
os.environ["ARFE_FWD_SKIP"] = "0"
