"""Bring-up probe of the fused NonLocal2D attention (arfe_nonlocal_attention_forward).

Each case runs in its own process (a trapped kernel poisons the CUDA context), prints the error
against an fp32 torch reference and against a reference whose operands are rounded to bf16 like
the kernel's, and a few structured inputs that localise a layout error:
  zero-theta  : P uniform           -> y = mean over positions of g   (P V product, V layout, O read-out)
  const-g     : g constant per chan -> y = that constant              (channel mapping)
  onehot-g    : g[q, d] = (q == d)  -> y[p, d] = P[p, d]              (Q K^T product, softmax)
Usage: python scripts/nonlocal_probe.py            (all cases)
       python scripts/nonlocal_probe.py --case i   (one case, in-process)
       python scripts/nonlocal_probe.py --time     (timing of the bench shape)
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

CASES = [
    # (B, D, H, W, layout, dtype, nsplit, kind, gain)
    (1, 64, 8, 8, "nchw", "f32", 1, "zero-theta", 1.0),
    (1, 64, 8, 8, "nchw", "f32", 1, "const-g", 1.0),
    (1, 64, 8, 8, "nchw", "f32", 1, "onehot-g", 1.0),
    (1, 64, 8, 8, "nchw", "f32", 1, "random", 1.0),
    (1, 64, 16, 8, "nchw", "f32", 1, "random", 1.0),
    (1, 64, 16, 16, "nchw", "f32", 1, "random", 1.0),
    (1, 256, 8, 8, "nchw", "f32", 1, "onehot-g", 1.0),
    (1, 256, 16, 16, "nchw", "f32", 1, "random", 1.0),
    (1, 256, 16, 16, "nhwc", "f32", 1, "random", 1.0),
    (2, 256, 25, 42, "nchw", "f32", 1, "random", 1.0),
    (2, 256, 25, 42, "nchw", "f32", 2, "random", 1.0),
    (2, 128, 25, 42, "nhwc", "bf16", 3, "random", 1.0),
    (2, 256, 50, 84, "nchw", "f32", 1, "random", 1.0),
    (2, 256, 50, 84, "nhwc", "f32", 2, "random", 3.0),   # peaky softmax: exercises the rescale path
    (2, 256, 50, 84, "nchw", "bf16", 2, "random", 1.0),
]


def make(case):
    import torch
    B, D, H, W, layout, dtype, nsplit, kind, gain = case
    g = torch.Generator(device="cuda").manual_seed(1234)
    HW = H * W
    sd = gain / D ** 0.25      # logits ~ N(0, gain^2)
    theta = torch.randn(B, D, H, W, generator=g, device="cuda") * sd
    phi = torch.randn(B, D, H, W, generator=g, device="cuda") * sd
    gx = torch.randn(B, D, H, W, generator=g, device="cuda")
    if kind == "zero-theta":
        theta.zero_()
    elif kind == "const-g":
        gx = torch.arange(D, device="cuda", dtype=torch.float32).view(1, D, 1, 1).expand(B, D, H, W).contiguous() / D
    elif kind == "onehot-g":
        gx = torch.zeros(B, D, HW, device="cuda")
        n = min(D, HW)
        idx = torch.arange(n, device="cuda")
        gx[:, idx, idx] = 1.0
        gx = gx.view(B, D, H, W)
    dt = torch.float32 if dtype == "f32" else torch.bfloat16
    ts = [t.to(dt) for t in (theta, phi, gx)]
    if layout == "nhwc":
        ts = [t.contiguous(memory_format=torch.channels_last) for t in ts]
    return ts, nsplit


def reference(theta, phi, gx, round_bf16):
    import torch
    B, D, H, W = theta.shape
    f = torch.float32
    th = theta.reshape(B, D, -1).permute(0, 2, 1).to(f)
    ph = phi.reshape(B, D, -1).to(f)
    g = gx.reshape(B, D, -1).permute(0, 2, 1).to(f)
    if round_bf16:
        th, ph, g = (t.to(torch.bfloat16).to(f) for t in (th, ph, g))
    p = torch.matmul(th, ph).softmax(dim=-1)
    y = torch.matmul(p, g)
    return y.permute(0, 2, 1).reshape(B, D, H, W)


def run_case(i):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    import arfe_b200 as A
    case = CASES[i]
    (theta, phi, gx), nsplit = make(case)
    y = A.nonlocal_attention(theta, phi, gx, 1.0, nsplit)
    torch.cuda.synchronize()
    ref = reference(theta, phi, gx, False)
    refb = reference(theta, phi, gx, True)
    yf = y.float()
    scale = ref.abs().max().item()
    e32 = (yf - ref).abs().max().item()
    eb = (yf - refb).abs().max().item()
    bad = torch.isnan(yf).sum().item()
    print(f"case {i:2d} {case}: max|y-ref32| {e32:.3e}  max|y-ref_bf16ops| {eb:.3e}  scale {scale:.3e}  nan {bad}",
          flush=True)
    if eb > 2e-2 * scale + 1e-3:
        d = (yf - refb).abs()
        b, c, h, w = [int(v) for v in torch.unravel_index(d.argmax(), d.shape)]
        print(f"    worst at b={b} c={c} h={h} w={w}: got {yf[b, c, h, w].item():.5f} want {refb[b, c, h, w].item():.5f}")
        row = yf[0, :8, 0, :8].cpu()
        print("    y[0,:8,0,:8]\n", row, "\n    ref\n", refb[0, :8, 0, :8].cpu(), flush=True)


def run_time():
    import torch
    import arfe_b200 as A
    torch.backends.cuda.matmul.allow_tf32 = False
    for case in [(2, 256, 50, 84, "nchw", "f32", 1, "random", 1.0), (2, 256, 50, 84, "nchw", "f32", 2, "random", 1.0),
                 (2, 256, 50, 84, "nhwc", "f32", 2, "random", 1.0), (2, 256, 50, 84, "nhwc", "bf16", 2, "random", 1.0),
                 (8, 256, 25, 42, "nchw", "f32", 2, "random", 1.0),
                 (2, 128, 50, 84, "nhwc", "bf16", 2, "random", 1.0), (2, 128, 50, 84, "nhwc", "bf16", 1, "random", 1.0)]:
        (theta, phi, gx), nsplit = make(case)
        for _ in range(3):
            A.nonlocal_attention(theta, phi, gx, 1.0, nsplit)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            A.nonlocal_attention(theta, phi, gx, 1.0, nsplit)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / n
        # the same call replayed as a CUDA graph: device time without the Python / ctypes launch path
        gr = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            A.nonlocal_attention(theta, phi, gx, 1.0, nsplit)
            with torch.cuda.graph(gr, stream=st):
                A.nonlocal_attention(theta, phi, gx, 1.0, nsplit)
        torch.cuda.current_stream().wait_stream(st)
        gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        tg = e0.elapsed_time(e1) / n
        e0.record()
        for _ in range(n):
            reference(theta, phi, gx, False)
        e1.record()
        torch.cuda.synchronize()
        tr = e0.elapsed_time(e1) / n
        B, D, H, W = theta.shape
        flops = 4.0 * B * (H * W) ** 2 * D
        print(f"{case}: fused {t * 1e3:.1f} us, as a graph {tg * 1e3:.1f} us ({flops / tg / 1e9:.1f} TFLOP/s)   "
              f"torch eager {tr * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    if "--case" in sys.argv:
        run_case(int(sys.argv[sys.argv.index("--case") + 1]))
    elif "--time" in sys.argv:
        run_time()
    else:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True,
                                   timeout=180)
                out = r.stdout.strip()
                print(out if out else f"case {i} {CASES[i]}: rc={r.returncode} {r.stderr.strip()[-400:]}", flush=True)
                if r.returncode != 0 and out:
                    print(f"    rc={r.returncode} {r.stderr.strip()[-300:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"case {i} {CASES[i]}: TIMEOUT", flush=True)
