"""Shared helpers for the parity tests."""
import numpy as np
import torch

STRIDES = (4, 8, 16, 32, 64)


def small_pyramid(oracle, batch=2, channels=16, img_h=192, img_w=320, seed=0,
                  strides=STRIDES):
    shapes = oracle.pyramid_shapes(img_h, img_w, strides)
    return oracle.synthetic_pyramid(batch, channels, shapes, seed=seed)


def edge_rois(img_w, img_h, batch):
    """RoIs the reference's callers can produce at the borders of validity."""
    r = [
        [0, 10.0, 10.0, 10.0, 10.0],                    # zero area
        [0, 0.0, 0.0, img_w - 1.0, img_h - 1.0],        # whole image
        [batch - 1, 0.0, 0.0, 3.0, 2.0],                # tiny at the corner
        [0, img_w - 5.0, img_h - 4.0, img_w - 1.0, img_h - 1.0],  # far corner
        [0, 5.0, 1.0, 8.0, img_h - 2.0],                # extreme tall
        [batch - 1, 2.0, 7.0, img_w - 3.0, 11.0],       # extreme wide
        [0, 20.5, 30.25, 76.5, 86.25],                  # exactly 56x56 -> level boundary
        [0, 16.0, 16.0, 128.0, 128.0],                  # 112x112 -> boundary of level 1
        [0, 0.0, 0.0, 0.5, 0.5],                        # sub-pixel
        [batch - 1, 100.0, 50.0, 101.0, 51.0],          # 1 px
    ]
    return torch.tensor(r, dtype=torch.float32)


def mixed_rois(oracle, K, img_w, img_h, batch, seed=0):
    smax = float(min(img_w, img_h)) * 0.9
    rois = oracle.synthetic_rois(K, img_w, img_h, batch, seed=seed, smin=6.0, smax=smax)
    return torch.cat([rois, edge_rois(img_w, img_h, batch)]).contiguous()


def assert_close_fp32(got, ref, what=""):
    """north_star tolerance for fp32: 1e-5 relative.  A 1e-6 absolute floor
    covers outputs that are sums cancelling to ~0 (features are N(0,1))."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = 1e-5 * ref.abs() + 1e-6
    bad = err > tol
    assert not bad.any(), (
        f"{what}: {int(bad.sum())}/{bad.numel()} elements off; max abs err "
        f"{float(err.max()):.3e}, max ref {float(ref.abs().max()):.3e}")


def assert_close_bf16(got, ref, what=""):
    """north_star tolerance for bf16 I/O: 1e-2 relative (to the tensor's scale
    for near-zero elements)."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = 1e-2 * ref.abs() + 1e-2 * float(ref.abs().mean() + 1e-12)
    bad = err > tol
    assert not bad.any(), (
        f"{what}: {int(bad.sum())}/{bad.numel()} elements off; max abs err "
        f"{float(err.max()):.3e}")
