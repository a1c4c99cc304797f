import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl
sys.path.insert(0, 'scripts')
import sweep
dev = torch.device("cuda:0")
ROI = ["roi_fuse_fwd", "rff_gate_fwd", "rff_gate_bwd", "roi_fuse_bwd"]
for dt in (torch.float32, torch.bfloat16):
    host = wl.host_inputs(2, 128, 256, dtype=dt, channels_last=True, out_size=14)
    st = wl.TrainStep(host, dev, regions=1)
    sweep.report(f"mask branch 14x14 x1 {dt}", st, sweep.time_ops(st, ROI[:1] + ROI[3:]), 2)
host = wl.host_inputs(2, 512, 256, channels_last=True, out_size=14)
st = wl.TrainStep(host, dev, regions=3)
sweep.report("14x14 x3 regions 2x512 f32", st, sweep.time_ops(st, ROI[:1] + ROI[3:]), 2)
