"""CPU suite: host-side logic that needs no GPU -- the registry swap by
subclassing, the gate's layout classification, the workload's RoI order and the
oracle's composition of the bench step on a tiny case."""
import torch


def test_accelerate_class_keeps_reference_methods():
    import arfe_b200 as A

    class RefHead(A.MultiRoIsBBoxHead):      # stands for the reference's registered class
        def loss(self):
            return "reference loss"

        def forward(self, x):
            raise RuntimeError("reference forward must be replaced")

    H = A.accelerate_class(RefHead, A.MultiBBoxHead, ("forward", "fuse"))
    assert H.__name__ == "RefHead" and issubclass(H, RefHead)
    assert H.forward is A.MultiBBoxHead.forward and H.fuse is A.MultiBBoxHead.fuse
    h = H(in_channels=8, fc_out_channels=16, num_classes=3)
    assert h.loss() == "reference loss"
    assert set(h.state_dict()) == set(RefHead(in_channels=8, fc_out_channels=16, num_classes=3).state_dict())
    assert A.accelerate_class(H, A.MultiBBoxHead, ("forward", "fuse")) is H   # idempotent
    assert A.register_into_mmdet() is False                                     # mmdet is absent here


def test_gate_layout_classification():
    from arfe_b200.functional import _gate_rows
    x = torch.randn(5, 48, 7, 7)
    assert _gate_rows(x[:, :16]) == ("nchw", 5, 16 * 49, 48 * 49)
    xc = x.contiguous(memory_format=torch.channels_last)
    assert _gate_rows(xc[:, :16]) == ("nhwc", 5 * 49, 16, 48)
    one = torch.randn(1, 16, 7, 7).contiguous(memory_format=torch.channels_last)
    assert _gate_rows(one) == ("nhwc", 49, 16, 16)                   # K == 1, channels-last
    assert _gate_rows(torch.randn(1, 48, 7, 7).contiguous(memory_format=torch.channels_last)[:, :16]) == \
        ("nhwc", 49, 16, 48)                                          # K == 1 slice of the cat tensor
    assert _gate_rows(torch.randn(1, 16, 7, 7)) == ("nchw", 1, 16 * 49, 16 * 49)
    assert _gate_rows(x[:, :, ::2]) is None                           # anything else: densified first


def test_roi_order_and_reference_step():
    from arfe_b200 import workload as wl
    from oracle import arfe_oracle as O
    r = wl.synthetic_rois(10, batch=2, seed=1)
    assert r[:, 0].tolist() == [0.0] * 5 + [1.0] * 5                   # image-major blocks (bbox2roi)
    ri = wl.synthetic_rois(10, batch=2, seed=1, order="interleaved")
    assert ri[:, 0].tolist() == [0.0, 1.0] * 5
    assert torch.equal(r[:, 1:], ri[:, 1:])
    assert torch.equal(ri, O.synthetic_rois(10, batch=2, seed=1))
    assert torch.equal(r, O.synthetic_rois(10, batch=2, seed=1, order="image_major"))
    host = wl.host_inputs(batch=2, rois_per_img=6, channels=8, img_h=96, img_w=128, smin=8.0, smax=80.0)
    full = O.reference_step(host)
    sub = O.reference_step(host, channels=[1, 6], dy_full=full["dy"])
    # RoIAlign and the gate are per channel: a channel subset reproduces the full run exactly
    for k in ("z", "d_ori", "d_ab"):
        assert torch.equal(sub[k], full[k][:, [1, 6]]), k
    assert torch.equal(sub["F"], torch.cat([full["F"][:, [1, 6]], full["F"][:, [9, 14]], full["F"][:, [17, 22]]], 1))
    for l in range(5):
        assert torch.equal(sub["dy"][l], full["dy"][l][:, [1, 6]])
        assert torch.equal(sub["dx"][l], full["dx"][l])
    assert full["dy"][4].abs().max() == 0                               # the extractor reads 4 levels
