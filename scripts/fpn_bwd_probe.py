"""The fused AR-FPN backward on the bench pyramid (ncu target / quick timing)."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L  # noqa: E402

dev = torch.device("cuda:0")
host = wl.host_inputs(2, 64, 256, channels_last=True, device=dev)
st = wl.TrainStep(host, dev)
st.step()
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    L.check(st.fpn_bwd(), "fpn_bwd")
b.record()
torch.cuda.synchronize()
print("fpn_bwd us", a.elapsed_time(b) / n * 1e3)
