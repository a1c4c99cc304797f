"""CUDA-event time of the RoI backward (bins ready: pull kernel + its two list kernels) under the
ARFE_BWD_SKIP knobs (1 no math, 2 no bin copies, 3 neither), for ARFE_PULL_NP = 1 and default."""
import os, subprocess, sys
import torch
sys.path.insert(0, os.getcwd())

def run():
    from arfe_b200 import workload as wl, _lib as L
    dev = torch.device("cuda:0")
    host = wl.host_inputs(2, 512, 256, channels_last=True)
    st = wl.TrainStep(host, dev)
    L.check(st.roi_fuse_fwd(), "f"); st.glue_before_roi_bwd()
    out = []
    for knob in ("0", "1", "2", "3"):
        os.environ["ARFE_BWD_SKIP"] = knob
        for _ in range(3):
            L.check(st.roi_fuse_bwd(), "b")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            L.check(st.roi_fuse_bwd(), "b")
        e1.record(); torch.cuda.synchronize()
        out.append("skip=%s %.1f us" % (knob, e0.elapsed_time(e1) / 30 * 1e3))
    print("ARFE_PULL_NP=%s" % os.environ.get("ARFE_PULL_NP", "default"), " | ".join(out))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for np_ in ("1", "0"):
            subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, ARFE_PULL_NP=np_), check=True)
