"""RoI forward + backward of the bench workload, a few times (ncu target)."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L
dev = torch.device("cuda:0")
host = wl.host_inputs(2, 512, 256, channels_last=True)
st = wl.TrainStep(host, dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    L.check(st.roi_fuse_fwd(), "f"); st.glue_before_roi_bwd(); L.check(st.roi_fuse_bwd(), "b")
torch.cuda.synchronize()
