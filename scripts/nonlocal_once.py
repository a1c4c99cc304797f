"""A few calls of the fused NonLocal2D attention at the bench shape (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import arfe_b200 as A

B, D, H, W = 2, 256, 50, 84
layout = sys.argv[1] if len(sys.argv) > 1 else "nhwc"
bf16 = len(sys.argv) > 3 and sys.argv[3] == "bf16"
nsplit = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "auto" else None
g = torch.Generator(device="cuda").manual_seed(0)
ts = [torch.randn(B, D, H, W, generator=g, device="cuda") * (0.25 if i < 2 else 1.0) for i in range(3)]
if bf16:
    ts = [t.bfloat16() for t in ts]
if layout == "nhwc":
    ts = [t.contiguous(memory_format=torch.channels_last) for t in ts]
for _ in range(4):
    y = A.nonlocal_attention(*ts, 1.0, nsplit)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
