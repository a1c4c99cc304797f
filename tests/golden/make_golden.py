"""Generate the committed golden vectors for the AR-RFF / AR-FPN path.

The reference ships no golden vectors or tests for this path (SURVEY.md
section 4), so these are minted here, in the build container, from the
REFERENCE'S OWN RoIAlign sources compiled unmodified (oracle/_ref, see
oracle/build_oracle.py) driven by the oracle's restatement of the Python
modules (which executes the same torch CPU ops the reference calls).

Run (needs /root/reference, i.e. this container only):
    python tests/golden/make_golden.py
Outputs small .npz files next to this script; they are committed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import arfe_oracle as O  # noqa: E402
from oracle import build_oracle  # noqa: E402
from util import STRIDES, mixed_rois, small_pyramid  # noqa: E402


def level_threshold_rois():
    rows = []
    for k in range(1, 5):
        side = np.float32(56.0 * 2.0 ** k)
        for _ in range(41):
            side = np.nextafter(side, np.float32(0), dtype=np.float32)
        for _ in range(80):
            rows.append([0.0, 0.0, 0.0, float(side), float(side)])
            side = np.nextafter(side, np.float32(1e9), dtype=np.float32)
    return torch.tensor(rows, dtype=torch.float32)


def main():
    build_oracle.build_c_oracle()
    assert build_oracle.build_reference_ext() is not None, "needs /root/reference"
    torch.manual_seed(0)

    # 1. AR-RFF extraction, forward + backward, via the compiled reference
    feats = small_pyramid(O, batch=2, channels=8, img_h=96, img_w=160, seed=21)
    rois = mixed_rois(O, 24, 160, 96, 2, seed=21)
    fo = [f.clone().requires_grad_(True) for f in feats]
    out = O.arrff_bbox_feats(fo, rois, list(STRIDES), backend="ref")
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(22))
    out.backward(g)
    boxes, lvls = O.region_boxes_and_levels(rois, 5)
    np.savez_compressed(
        os.path.join(HERE, "arrff_small.npz"),
        rois=rois.numpy(), boxes=boxes.numpy(), lvls=lvls.numpy().astype(np.int32),
        out=out.detach().numpy(), grad_out=g.numpy(),
        **{f"feat{l}": feats[l].numpy() for l in range(5)},
        **{f"dfeat{l}": (fo[l].grad if fo[l].grad is not None else torch.zeros_like(feats[l])).numpy()
           for l in range(5)})

    # 2. Level map around every threshold (torch CPU in this container)
    tr = level_threshold_rois()
    np.savez_compressed(os.path.join(HERE, "level_thresholds.npz"),
                        rois=tr.numpy(), lvls=O.map_roi_levels(tr, 5).numpy().astype(np.int32))

    # 3. RoIAlign operator (reference gradcheck.py:11-30 sizes) fwd/bwd
    gen = torch.Generator().manual_seed(23)
    feat = torch.randn(2, 16, 15, 15, generator=gen)
    r = torch.rand(20, 4, generator=gen) * 60.0
    r[:, 2:] += 60.0
    b = torch.randint(0, 2, (20, 1), generator=gen).float()
    r = torch.cat([b, r], dim=1)
    res = {}
    for sn in (0, 2):
        o = O.roi_align_forward(feat, r, 3, 1 / 8, sn, "ref")
        gg = torch.randn(o.shape, generator=gen)
        res[f"out_sn{sn}"] = o.numpy()
        res[f"gout_sn{sn}"] = gg.numpy()
        res[f"gin_sn{sn}"] = O.roi_align_backward(gg, r, 3, 1 / 8, feat.shape, sn, "ref").numpy()
    np.savez_compressed(os.path.join(HERE, "roi_align_op.npz"), feat=feat.numpy(),
                        rois=r.numpy(), **res)

    # 4. AR-FPN gather + apply (torch CPU ops = the reference arithmetic)
    shapes = [(24, 40), (12, 20), (6, 10), (3, 5), (2, 3)]
    xs = O.synthetic_pyramid(2, 8, shapes, seed=24)
    gen = torch.Generator().manual_seed(25)
    bsf = torch.randn(2, 8, 6, 10, generator=gen)
    g1 = [torch.randn(2, 1, h, w, generator=gen) for h, w in shapes]
    g2 = [torch.randn(2, 1, h, w, generator=gen) for h, w in shapes]
    gathered = O.wfpn_gather(xs, 2)
    outs = O.wfpn_apply(xs, bsf, g1, g2)
    np.savez_compressed(
        os.path.join(HERE, "arfpn_small.npz"), bsf=bsf.numpy(), gathered=gathered.numpy(),
        **{f"x{l}": xs[l].numpy() for l in range(5)},
        **{f"g1_{l}": g1[l].numpy() for l in range(5)},
        **{f"g2_{l}": g2[l].numpy() for l in range(5)},
        **{f"out{l}": outs[l].numpy() for l in range(5)})
    # 5. NMS through the reference's own nms_ext (CPU build, compiled unmodified)
    assert build_oracle.build_reference_nms_ext() is not None, "needs /root/reference"
    gen = torch.Generator().manual_seed(26)
    n = 600
    ctr = torch.rand(n, 2, generator=gen) * 300
    wh = torch.rand(n, 2, generator=gen) * 80 + 2
    dets = torch.cat([ctr - wh / 2, ctr + wh / 2, torch.rand(n, 1, generator=gen)], 1).contiguous()
    res = {"dets": dets.numpy()}
    for thr in (0.3, 0.5, 0.7):
        res[f"keep_{int(thr * 10)}"] = O.nms(dets, thr, "ref").numpy()
    np.savez_compressed(os.path.join(HERE, "nms_small.npz"), **res)

    # 6. NonLocal2D attention (restated non_local.py:65-101; torch CPU executes the arithmetic)
    gen = torch.Generator().manual_seed(27)
    th = torch.randn(2, 64, 9, 11, generator=gen) * 0.4
    ph = torch.randn(2, 64, 9, 11, generator=gen) * 0.4
    gx = torch.randn(2, 64, 9, 11, generator=gen)
    np.savez_compressed(os.path.join(HERE, "nonlocal_small.npz"), theta=th.numpy(), phi=ph.numpy(), g=gx.numpy(),
                        y=O.nonlocal_attention(th, ph, gx).numpy(),
                        y_scaled=O.nonlocal_attention(th, ph, gx, use_scale=True).numpy())

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
