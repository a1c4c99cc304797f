"""The fused gate convolutions (arfe_fpn_gate_conv_forward) against cuDNN's 2 x 5 Conv2d(C, 1, 3) on the
bench pyramid, channels-last, fp32 true precision (TF32 off) and TF32."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.getcwd())
import arfe_b200 as A  # noqa: E402
from arfe_b200 import workload as wl  # noqa: E402

dev = torch.device("cuda:0")
shapes = wl.pyramid_shapes(800, 1344)
B, C = 2, 256
xs = [torch.randn(B, C, h, w, device=dev).contiguous(memory_format=torch.channels_last) for h, w in shapes]
w1 = [torch.randn(1, C, 3, 3, device=dev) * 0.05 for _ in shapes]
w2 = [torch.randn(1, C, 3, 3, device=dev) * 0.05 for _ in shapes]
b1 = [torch.randn(1, device=dev) for _ in shapes]
b2 = [torch.randn(1, device=dev) for _ in shapes]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def cudnn():
    return [F.conv2d(x, w, b, padding=1) for x, w, b in zip(xs, w1, b1)] + \
           [F.conv2d(x, w, b, padding=1) for x, w, b in zip(xs, w2, b2)]


with torch.no_grad():
    t_ours = timeit(lambda: A.fpn_gate_conv(xs, w1, b1, w2, b2))
    torch.backends.cudnn.allow_tf32 = False
    t_fp32 = timeit(cudnn)
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timeit(cudnn)

# backward: d x, d weights, d biases of all ten convolutions
xg = [x.clone().requires_grad_(True) for x in xs]
pw = [t.clone().requires_grad_(True) for t in w1 + w2 + b1 + b2]
n = len(shapes)


def bwd_ours():
    g1, g2 = A.fpn_gate_conv(xg, pw[:n], pw[2 * n:3 * n], pw[n:2 * n], pw[3 * n:])
    torch.autograd.backward(list(g1) + list(g2), [torch.ones_like(t) for t in list(g1) + list(g2)])


def bwd_cudnn():
    outs = [F.conv2d(x, w, b, padding=1) for x, w, b in zip(xg, pw[:n], pw[2 * n:3 * n])] + \
           [F.conv2d(x, w, b, padding=1) for x, w, b in zip(xg, pw[n:2 * n], pw[3 * n:])]
    torch.autograd.backward(outs, [torch.ones_like(t) for t in outs])


tb_ours = timeit(bwd_ours, 10)
torch.backends.cudnn.allow_tf32 = False
tb_fp32 = timeit(bwd_cudnn, 5)
torch.backends.cudnn.allow_tf32 = True
tb_tf32 = timeit(bwd_cudnn, 5)
print(f"forward + backward: fused {tb_ours:7.1f} us | cuDNN fp32 {tb_fp32:7.1f} us, tf32 {tb_tf32:7.1f} us")
pyr = sum(B * C * h * w * 4 for h, w in shapes)
print(f"fused gate convs {t_ours:7.1f} us ({pyr / t_ours / 1e3:6.0f} GB/s of pyramid) | cuDNN 10 convs fp32 {t_fp32:7.1f} us, tf32 {t_tf32:7.1f} us")
