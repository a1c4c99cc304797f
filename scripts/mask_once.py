import sys, os, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L
dev = torch.device("cuda:0")
host = wl.host_inputs(2, 128, 256, channels_last=True, out_size=14)
st = wl.TrainStep(host, dev, regions=1)
for _ in range(3):
    L.check(st.roi_fuse_fwd(), "f"); st.glue_before_roi_bwd(); L.check(st.roi_fuse_bwd(), "b")
torch.cuda.synchronize()
