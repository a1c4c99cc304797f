"""the NonLocal2D refine (mmdet/ops/non_local.py:65-69, :78-101,
called at necks/wfpn_dual_spatial.py:115) as one fused tensor-core attention.

Tolerances.  The kernel rounds theta / phi / g (and the softmax weights) to bf16 and does
everything else in fp32, so against the fp32 reference it is held to north_star's bf16 bound,
|delta| <= 1e-2 * max|ref|; against the oracle evaluated on the SAME bf16-rounded operands (only
the rounding of the softmax weights and the summation order differ) to 3e-3 * max|ref|."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mk(B, D, H, W, gain=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = gain / D ** 0.25          # logits ~ N(0, gain^2)
    return (torch.randn(B, D, H, W, generator=g) * sd, torch.randn(B, D, H, W, generator=g) * sd,
            torch.randn(B, D, H, W, generator=g))


def _dev(ts, cuda, layout, dtype):
    out = [t.to(cuda).to(dtype) for t in ts]
    if layout == "nhwc":
        out = [t.contiguous(memory_format=torch.channels_last) for t in out]
    return out


def _check(y, ts, oracle, use_scale=False, tol32=1e-2, tolb=3e-3):
    ref = oracle.nonlocal_attention(*ts, use_scale=use_scale)
    refb = oracle.nonlocal_attention(*ts, use_scale=use_scale, round_operands=torch.bfloat16)
    got = y.detach().float().cpu()
    assert torch.isfinite(got).all()
    s = ref.abs().max()
    assert (got - ref).abs().max() <= tol32 * s, ((got - ref).abs().max() / s)
    assert (got - refb).abs().max() <= tolb * s, ((got - refb).abs().max() / s)


@pytest.mark.parametrize("B,D,H,W,layout,dtype,nsplit", [
    (1, 64, 8, 8, "nchw", torch.float32, 1),       # one key step
    (1, 64, 7, 11, "nhwc", torch.float32, 1),      # 77 positions: ragged query block and key step
    (2, 128, 13, 21, "nchw", torch.float32, 2),    # 273: the coarsest level of the bench pyramid
    (2, 256, 25, 42, "nchw", torch.float32, 1),    # RetinaNet refine level (BASELINE config 2)
    (2, 256, 25, 42, "nhwc", torch.bfloat16, 3),
    (1, 256, 16, 8, "nhwc", torch.float32, 2),     # exactly one 128-row query block, two key steps
    (3, 64, 20, 13, "nchw", torch.bfloat16, 4),
    (1, 64, 3, 5, "nchw", torch.float32, 1),       # 15 positions (the coarsest level): half a key step is empty
    (2, 64, 3, 5, "nhwc", torch.bfloat16, 1),      # the same through the tensor maps (box larger than the tensor)
    (1, 64, 7, 11, "nhwc", torch.bfloat16, 1),     # tensor-map path: one ragged tile (zero fill past HW)
    (2, 128, 9, 15, "nhwc", torch.bfloat16, 2),    # tensor-map path: 135 positions, key range split
])
def test_attention_matches_oracle(oracle, cuda, B, D, H, W, layout, dtype, nsplit):
    import arfe_b200 as A
    ts = _mk(B, D, H, W, seed=B * 100 + D + H)
    if dtype == torch.bfloat16:
        ts = tuple(t.to(dtype).float() for t in ts)          # the oracle sees what the kernel sees
    y = A.nonlocal_attention(*_dev(ts, cuda, layout, dtype), 1.0, nsplit)
    assert y.dtype == dtype and y.shape == ts[0].shape
    if layout == "nhwc":
        assert y.is_contiguous(memory_format=torch.channels_last)
    _check(y, ts, oracle, tol32=1e-2 if dtype == torch.float32 else 1.5e-2,
           tolb=3e-3 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_attention_bench_shape(oracle, cuda, layout):
    """BASELINE configs[1] refine level: 2 images, 256 channels, 50 x 84 = 4200 positions (the
    shape the kernel is timed on), default key-range split."""
    import arfe_b200 as A
    ts = _mk(2, 256, 50, 84, seed=7)
    y = A.nonlocal_attention(*_dev(ts, cuda, layout, torch.float32))
    _check(y, ts, oracle)


def test_attention_golden(cuda):
    import arfe_b200 as A
    z = np.load(os.path.join(GOLD, "nonlocal_small.npz"))
    ts = [torch.from_numpy(z[k]) for k in ("theta", "phi", "g")]
    for key, scale in (("y", 1.0), ("y_scaled", 1.0 / 8.0)):
        want = torch.from_numpy(z[key])
        y = A.nonlocal_attention(*_dev(ts, cuda, "nchw", torch.float32), scale).cpu()
        assert (y - want).abs().max() <= 1e-2 * want.abs().max()


def test_attention_structured_inputs(cuda):
    """Inputs whose answer is known in closed form: zero theta -> uniform weights -> the mean of g
    over the positions; one-hot g -> the softmax weights themselves."""
    import arfe_b200 as A
    B, D, H, W = 1, 64, 12, 10
    HW = H * W
    th, ph, gx = _mk(B, D, H, W, seed=3)
    y = A.nonlocal_attention(*_dev((torch.zeros_like(th), ph, gx), cuda, "nchw", torch.float32)).cpu()
    mean = gx.to(torch.bfloat16).float().reshape(B, D, HW).mean(-1)
    assert (y.reshape(B, D, HW) - mean[..., None]).abs().max() <= 2e-3
    onehot = torch.zeros(B, D, HW)
    onehot[:, torch.arange(D), torch.arange(D)] = 1.0
    y = A.nonlocal_attention(*_dev((th, ph, onehot.view(B, D, H, W)), cuda, "nchw", torch.float32)).cpu()
    q = th.to(torch.bfloat16).float().reshape(B, D, HW).permute(0, 2, 1)
    k = ph.to(torch.bfloat16).float().reshape(B, D, HW)
    p = torch.matmul(q, k).softmax(-1)                        # [B, HW, HW]
    assert (y.reshape(B, D, HW).permute(0, 2, 1) - p[:, :, :D]).abs().max() <= 4e-3 * p.max()


def test_attention_growing_maximum(oracle, cuda):
    """Keys ordered so that every row's maximum keeps growing by far more than the rescale
    threshold (2^8) from key step to key step: exercises the in-TMEM correction of O."""
    import arfe_b200 as A
    B, D, H, W = 1, 64, 16, 24                                # 384 positions = 6 key steps
    HW = H * W
    g = torch.Generator().manual_seed(11)
    th = torch.ones(B, D, HW) * 0.5 + torch.randn(B, D, HW, generator=g) * 0.05
    ramp = torch.linspace(0.1, 3.0, HW).view(1, 1, HW)       # logits from ~3 to ~100
    ph = torch.ones(B, D, HW) * ramp
    gx = torch.randn(B, D, HW, generator=g)
    ts = tuple(t.view(B, D, H, W).contiguous() for t in (th, ph, gx))
    for nsplit in (1, 2):
        y = A.nonlocal_attention(*_dev(ts, cuda, "nchw", torch.float32), 1.0, nsplit)
        refb = oracle.nonlocal_attention(*ts, round_operands=torch.bfloat16)
        got = y.float().cpu()
        assert torch.isfinite(got).all()
        assert (got - refb).abs().max() <= 1e-2 * refb.abs().max()


def test_attention_use_scale(oracle, cuda):
    import arfe_b200 as A
    ts = _mk(2, 64, 10, 10, gain=4.0, seed=5)
    y = A.nonlocal_attention(*_dev(ts, cuda, "nchw", torch.float32), 1.0 / 64 ** 0.5)
    _check(y, ts, oracle, use_scale=True)


def test_attention_split_agrees(cuda):
    """The key-range split changes the order of the partial sums and the reference maximum the
    bf16 softmax weights are rounded at: results agree to the bf16-weight rounding."""
    import arfe_b200 as A
    ts = _dev(_mk(2, 256, 25, 42, seed=9), cuda, "nhwc", torch.float32)
    y1 = A.nonlocal_attention(*ts, 1.0, 1)
    for ns in (2, 3, 4):
        yn = A.nonlocal_attention(*ts, 1.0, ns)
        assert (y1 - yn).abs().max() <= 1e-2 * y1.abs().max(), float((y1 - yn).abs().max() / y1.abs().max())


def test_attention_rejects(cuda):
    import arfe_b200 as A
    t = torch.randn(1, 96, 4, 4, device=cuda)
    with pytest.raises(RuntimeError):
        A.nonlocal_attention(t, t, t)                          # inter_channels not 64 / 128 / 256
    c = torch.randn(1, 64, 4, 4)
    with pytest.raises(RuntimeError):
        A.nonlocal_attention(c, c, c)                          # no CPU path
    t = torch.randn(1, 64, 4, 4, device=cuda)
    with pytest.raises(RuntimeError):
        A.nonlocal_attention(t, t, t, -1.0)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_nonlocal_module_forward_backward(oracle, cuda, layout):
    """NonLocal2D(fused_attention=True) against the oracle module with trained-like weights:
    forward to the bf16 bound; backward (recomputed from the saved operands with library matmuls)
    of input and parameters."""
    import arfe_b200 as A
    torch.manual_seed(21)
    C = 64
    ref = oracle.NonLocal2D(C)
    for p in ref.parameters():
        torch.nn.init.normal_(p, 0, 0.08)
    mod = A.NonLocal2D(C, reduction=1, use_scale=False)
    mod.load_state_dict(ref.state_dict())
    mod.fused_attention = True
    mod.to(cuda)
    x = torch.randn(2, C, 14, 18)
    xr = x.clone().requires_grad_(True)
    out_ref = ref(xr)
    w = torch.randn_like(out_ref)
    (out_ref * w).sum().backward()
    xg = x.to(cuda)
    if layout == "nhwc":
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    out = mod(xg)
    (out * w.to(cuda)).sum().backward()
    s = (out_ref - x).abs().max()                              # scale of the attention branch
    assert (out.detach().cpu() - out_ref.detach()).abs().max() <= 1e-2 * s
    assert (xg.grad.cpu() - xr.grad).abs().max() <= 2e-2 * xr.grad.abs().max()
    # the gradient of phi's bias is zero in exact arithmetic (a per-row constant shift of the logits
    # leaves the softmax unchanged): what comes out is the rounding noise of d S, which the backward
    # holds in bf16 like the forward holds the weights -- hence the floor relative to the largest gradient
    floor = 2e-3 * max(float(q.grad.abs().max()) for q in ref.parameters())
    for (n, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
        assert (p.grad.cpu() - q.grad).abs().max() <= 2e-2 * q.grad.abs().max() + floor, n


def test_nonlocal_module_policy(cuda):
    """'auto' keeps an fp32 module on the reference's fp32 arithmetic and fuses bf16 activations;
    optimize_detector(fused_attention=True) opts an fp32 model in."""
    import arfe_b200 as A
    mod = A.NonLocal2D(64, reduction=1, use_scale=False).to(cuda)
    x = torch.randn(1, 64, 6, 6, device=cuda)
    assert mod.fused_attention == 'auto' and not mod._use_fused(x) and mod._use_fused(x.bfloat16())
    prec = torch.get_float32_matmul_precision()
    try:  # PyTorch's own switch for "bf16 is fine inside float32 matmuls"
        torch.set_float32_matmul_precision('medium')
        assert mod._use_fused(x)
    finally:
        torch.set_float32_matmul_precision(prec)
    assert not mod._use_fused(x)
    neck = A.WFPNDualSpatial(64, 5)
    A.optimize_detector(neck, fused_attention=True)
    assert neck.refine.fused_attention is True and neck.refine._use_fused(x)
    m2 = A.NonLocal2D(96, reduction=1).to(cuda)               # constructing is fine ...
    m2.fused_attention = True
    with pytest.raises(RuntimeError):
        m2(torch.randn(1, 96, 4, 4, device=cuda))              # ... running the fused path is refused


def test_neck_bf16_uses_fused_attention_and_gate_convs(oracle, cuda):
    """WFPNDualSpatial on bf16 channels-last maps takes every fused kernel of the neck -- gather,
    tensor-core attention ('auto' policy), one-pass gate convolutions, gated residual -- and stays
    within the bf16 bound of the fp32 oracle module with the same weights, forward and input gradients."""
    import arfe_b200 as A
    torch.manual_seed(1)
    C, shapes = 64, [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5)]
    ref_m = oracle.WFPNDualSpatial(C, 5)
    ref_m.init_weights()
    for p in ref_m.refine.parameters():
        torch.nn.init.normal_(p, 0, 0.05)
    m = A.WFPNDualSpatial(C, 5)
    m.load_state_dict(ref_m.state_dict())
    m = m.to(cuda).to(torch.bfloat16).to(memory_format=torch.channels_last)
    xs = [x.bfloat16().float() for x in oracle.synthetic_pyramid(2, C, shapes, seed=6)]
    calls = []
    real = A.neck.nonlocal_attention
    A.neck.nonlocal_attention = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
    try:
        xg = [x.to(cuda, torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True) for x in xs]
        out = m(xg)
        gs = [torch.randn(x.shape, generator=torch.Generator().manual_seed(40 + i)) for i, x in enumerate(xs)]
        torch.autograd.backward(list(out), [g.to(cuda, torch.bfloat16) for g in gs])
    finally:
        A.neck.nonlocal_attention = real
    assert calls, "the bf16 neck did not take the fused attention"
    xo = [x.clone().requires_grad_(True) for x in xs]
    ref = ref_m(xo)
    torch.autograd.backward(list(ref), [g.bfloat16().float() for g in gs])
    for l in range(5):
        assert out[l].dtype == torch.bfloat16
        err = (out[l].float().cpu() - ref[l].detach()).abs().max()
        assert float(err) <= 2e-2 * float(ref[l].abs().max()), (l, float(err))
        # gradients: the gates are tanh(relu(conv)); where a pre-activation sits within bf16 rounding of
        # zero the relu's derivative flips and that pixel's gradient legitimately differs by O(1), so the
        # bound is on the relative L2 error and on all but 0.5 % of the elements, not on the maximum
        d = (xg[l].grad.float().cpu() - xo[l].grad).abs()
        scale = float(xo[l].grad.abs().max())
        assert float(d.norm() / xo[l].grad.norm()) <= 3e-2, (l, float(d.norm() / xo[l].grad.norm()))
        assert float((d > 3e-2 * scale).float().mean()) <= 5e-3, (l, float((d > 3e-2 * scale).float().mean()))


def test_backward_row_pass(cuda):
    """arfe_nonlocal_backward_rows: P = softmax(scale S) and dS = scale P (dP - sum P dP) per row, as bf16."""
    from arfe_b200 import _lib as L
    g = torch.Generator().manual_seed(13)
    rows, n, scale = 37, 4200, 0.7
    S = (torch.randn(rows, n, generator=g) * 3).to(cuda)
    dP = torch.randn(rows, n, generator=g).to(cuda)
    Pb = torch.empty(rows, n, dtype=torch.bfloat16, device=cuda)
    dSb = torch.empty_like(Pb)
    L.check(L.lib().arfe_nonlocal_backward_rows(S.data_ptr(), dP.data_ptr(), Pb.data_ptr(), dSb.data_ptr(), rows, n,
                                                scale, L.stream_ptr(cuda)), "rows")
    P = (S.double() * scale).softmax(-1)
    dS = scale * P * (dP.double() - (P * dP.double()).sum(-1, keepdim=True))
    assert float((Pb.double() - P).abs().max()) <= 4e-3 * float(P.max())
    assert float((dSb.double() - dS).abs().max()) <= 4e-3 * float(dS.abs().max()) + 1e-9
