"""Does the forward ring kernel gain from spatially ordered RoIs?  Times the
forward of the bench workload with the RoIs as generated (random order) and
sorted by (image, level of the original box, y, x)."""
import os, sys, math, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L
dev = torch.device("cuda:0")
def run(host, tag, n=30):
    st = wl.TrainStep(host, dev)
    for _ in range(5): L.check(st.roi_fuse_fwd(), "f")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): L.check(st.roi_fuse_fwd(), "f")
    b.record(); torch.cuda.synchronize()
    print(f"{tag:28s}: {a.elapsed_time(b) / n * 1e3:7.1f} us")
host = wl.host_inputs(2, 512, 256, channels_last=True)
run(host, "random order")
r = host["rois"]
scale = ((r[:, 3] - r[:, 1]) * (r[:, 4] - r[:, 2])).clamp(min=1).sqrt()
lvl = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(0, 4)
key = ((r[:, 0] * 8 + lvl) * 4096 + ((r[:, 2] + r[:, 4]) / 2 / 16).floor()) * 4096 + (r[:, 1] + r[:, 3]) / 2
host2 = dict(host); host2["rois"] = r[torch.argsort(key)].contiguous()
run(host2, "sorted (image, level, y, x)")
