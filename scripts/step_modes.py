"""Step time of the bench workload: plain launches, launches with per-op events (what bench.py
times), and one CUDA-graph replay per step (capture of the same 14 launches + the second stream)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from arfe_b200 import workload as wl, _lib as L
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
host = wl.host_inputs(2, 512, 256, seed=0, pin=True, channels_last=True)
st = wl.TrainStep(host, dev)
N = 100

def timeit(fn, n=N):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

print("plain step          %.1f us" % timeit(st.step))
def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); L.check(fn(), name); b.record()
print("per-op events       %.1f us" % timeit(lambda: st.step(timed)))
if hasattr(st, "capture"):
    st.capture()
    print("graph replay        %.1f us" % timeit(st.step))
