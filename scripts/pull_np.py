"""RoI backward of the bench workload with one / two producer warps in the pull kernel
(ARFE_PULL_NP=1 forces one): CUDA-event time of the pull call, and bitwise agreement."""
import os, subprocess, sys
import torch
sys.path.insert(0, os.getcwd())

def run():
    from arfe_b200 import workload as wl, _lib as L
    dev = torch.device("cuda:0")
    host = wl.host_inputs(2, 512, 256, channels_last=True)
    st = wl.TrainStep(host, dev)
    L.check(st.roi_fuse_fwd(), "f"); st.glue_before_roi_bwd()
    for _ in range(5):
        L.check(st.roi_fuse_bwd(), "b")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        L.check(st.roi_fuse_bwd(), "b")
    e1.record(); torch.cuda.synchronize()
    import hashlib
    h = hashlib.sha256()
    for t in st.dy:
        h.update(t.detach().cpu().numpy().tobytes())
    print(os.environ.get("ARFE_PULL_NP", "default"), "roi_fuse_bwd ms", e0.elapsed_time(e1) / 50, h.hexdigest()[:16])

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for np_ in ("1", "0"):
            env = dict(os.environ, ARFE_PULL_NP=np_)
            subprocess.run([sys.executable, __file__, "x"], env=env, check=True)
