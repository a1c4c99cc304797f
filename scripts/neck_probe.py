"""Module-level time of the AR-FPN neck (WFPNDualSpatial) on the configs[1] pyramid, channels-last:
forward and forward + backward, with the refine block's attention through the library (the fp32
default) and through the fused tensor-core kernel, fp32 and bf16."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import arfe_b200 as A  # noqa: E402
from arfe_b200 import workload as wl  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
shapes = wl.pyramid_shapes(800, 1344)
B, C = 2, 256


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for dt, fused in ((torch.float32, False), (torch.float32, True), (torch.bfloat16, 'auto')):
    m = A.WFPNDualSpatial(C, 5).to(dev).to(dt).to(memory_format=torch.channels_last)
    m.init_weights()
    torch.nn.init.normal_(m.refine.conv_out.conv.weight, 0, 0.02)
    m.refine.fused_attention = fused
    xs = [torch.randn(B, C, h, w, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last).requires_grad_(True)
          for h, w in shapes]
    gs = [torch.randn_like(x) for x in xs]

    def fwd():
        with torch.no_grad():
            return m(xs)

    def fwdbwd():
        out = m(xs)
        torch.autograd.backward(list(out), gs)

    print(f"{str(dt)[6:]:9s} fused_attention={fused!s:5s}  forward {timeit(fwd):7.3f} ms   forward+backward {timeit(fwdbwd, 5):7.3f} ms",
          flush=True)
