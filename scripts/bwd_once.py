"""RoI backward of the bench workload under each ARFE_BWD_SKIP knob given on the command line (ncu target)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L
dev = torch.device("cuda:0")
host = wl.host_inputs(2, 512, 256, channels_last=True)
st = wl.TrainStep(host, dev)
L.check(st.roi_fuse_fwd(), "f"); st.glue_before_roi_bwd()
for knob in sys.argv[1:]:
    os.environ["ARFE_BWD_SKIP"] = knob
    for _ in range(2):
        L.check(st.roi_fuse_bwd(), "b")
    torch.cuda.synchronize()
