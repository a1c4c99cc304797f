"""CPU suite: the C-ABI library loads and exports every symbol that
include/arfe_b200.h declares; argument errors are reported without a GPU;
the Python surface fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "arfe_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arfe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from arfe_b200 import _lib, build
    build.build()
    names = _declared()
    assert len(names) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/arfe_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert _lib.lib().arfe_version() == 100


def test_argument_errors_need_no_gpu():
    from arfe_b200 import _lib as L
    lib = L.lib()
    H, W, s = L.int_array([8]), L.int_array([8]), L.float_array([0.25])
    # regions must be 1 or 3
    rc = lib.arfe_roi_fuse_forward(None, H, W, s, 1, 1, 4, None, 0, 2, 1.0, 7, 7, 0, 56.0,
                                   0, 0, 0, None, None, None, None)
    assert rc == -3 and b"regions" in lib.arfe_last_error()
    # pooled size out of range
    rc = lib.arfe_roi_fuse_forward(None, H, W, s, 1, 1, 4, None, 0, 3, 1.0, 64, 7, 0, 56.0,
                                   0, 0, 0, None, None, None, None)
    assert rc == -2
    # K = 0 is a no-op success (reference returns early, roi_align_kernel_v2.cu:293-296)
    rc = lib.arfe_roi_fuse_forward(None, H, W, s, 1, 1, 4, None, 0, 3, 1.0, 7, 7, 0, 56.0,
                                   0, 0, 0, None, None, None, None)
    assert rc == 0
    # NULL features with K > 0
    rois = torch.zeros(1, 5)
    rc = lib.arfe_roi_fuse_forward(None, H, W, s, 1, 1, 4, rois.data_ptr(), 1, 3, 1.0, 7, 7,
                                   0, 56.0, 0, 0, 0, None, None, None, None)
    assert rc == -1
    # channels-last output needs channels-last features
    rc = lib.arfe_roi_fuse_forward(None, H, W, s, 1, 1, 4, rois.data_ptr(), 1, 3, 1.0, 7, 7,
                                   0, 56.0, 0, 0, 1, None, None, None, None)
    assert rc in (-1, -5)
    H5 = L.int_array([200, 100, 50, 25, 13])
    W5 = L.int_array([336, 168, 84, 42, 21])
    assert lib.arfe_roi_fuse_pull_workspace_bytes(1024, 3, 5, 2, H5, W5) > 0
    assert lib.arfe_roi_fuse_pull_workspace_bytes(0, 3, 5, 2, H5, W5) == 0
    assert lib.arfe_roi_fuse_pull_workspace_bytes(1024, 3, 5, 2, H5, None) == 0
    assert lib.arfe_roi_plan_bytes(1024, 3, 5, 2, H5, W5) == \
        lib.arfe_roi_fuse_pull_workspace_bytes(1024, 3, 5, 2, H5, W5)
    # the planned forward validates its arguments before touching the device
    rc = lib.arfe_roi_fuse_forward_plan(None, H, W, s, 1, 1, 4, rois.data_ptr(), 1, 3, 1.0, 7, 7,
                                        0, 56.0, 0, None, None, None, None, 0, 0, None)
    assert rc == -1
    # aligned=False is the legacy path
    rc = lib.arfe_roi_align_forward(None, None, 0.25, 7, 7, 0, 0, 1, 4, 8, 8, 0, 0, 0, None, None)
    assert rc == -5
    rc = lib.arfe_rff_gate_forward(None, 10, None, None, None, 1, 20, 0, None)
    assert rc == -2  # stride < n_per_roi
    # NonLocal2D attention: channel counts outside the tensor-core tiles, bad splits, NULLs
    assert lib.arfe_nonlocal_workspace_bytes(2, 4200, 256, 2) > 2 * 4200 * 256 * 2 * 3
    assert lib.arfe_nonlocal_workspace_bytes(2, 4200, 96, 2) == 0
    rc = lib.arfe_nonlocal_attention_forward(None, None, None, None, 1, 64, 96, 0, 0, 1.0, 1, None, 0, None)
    assert rc == -5 and b"inter_channels" in lib.arfe_last_error()
    rc = lib.arfe_nonlocal_attention_forward(None, None, None, None, 1, 64, 64, 0, 0, 1.0, 2, None, 0, None)
    assert rc == -2  # more key slices than key tiles
    rc = lib.arfe_nonlocal_attention_forward(None, None, None, None, 1, 64, 64, 0, 0, -1.0, 1, None, 0, None)
    assert rc == -2  # scale must be positive
    rc = lib.arfe_nonlocal_attention_forward(None, None, None, None, 1, 64, 64, 0, 0, 1.0, 1, None, 0, None)
    assert rc == -1  # NULL tensors
    assert lib.arfe_nonlocal_attention_forward(None, None, None, None, 0, 64, 64, 0, 0, 1.0, 1, None, 0, None) == 0


def test_python_surface_refuses_cpu_tensors():
    import arfe_b200 as A
    ext = A.SingleRoIExtractor(dict(type='RoIAlign', out_size=7, sample_num=0), 8, [4, 8])
    feats = [torch.zeros(1, 8, 16, 16), torch.zeros(1, 8, 8, 8)]
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ext(feats, torch.zeros(2, 5))
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        A.WFPNDualSpatial(8, 2, refine_level=1)(feats)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        A.rff_gate(torch.zeros(2, 8, 7, 7), torch.zeros(2, 8, 7, 7), torch.zeros(2, 8, 7, 7))


def test_module_surface_and_parameter_names():
    """Checkpoint compatibility: parameter names."""
    import arfe_b200 as A
    neck = A.WFPNDualSpatial(in_channels=16, num_levels=5)
    keys = set(neck.state_dict())
    for i in range(5):
        for n in ("reduce_convs", "reduce_convs2"):
            assert f"{n}.{i}.conv.weight" in keys and f"{n}.{i}.conv.bias" in keys
    for n in ("g", "theta", "phi", "conv_out"):
        assert f"refine.{n}.conv.weight" in keys
    neck.init_weights()
    head = A.MultiRoIsBBoxHead(in_channels=16, fc_out_channels=32, roi_feat_size=7, num_classes=3)
    hk = set(head.state_dict())
    for n in ("hh_conv.conv.weight", "wh_conv.conv.weight", "final_conv.conv.bias",
              "shared_fcs.0.weight", "shared_fcs.1.bias", "fc_cls.weight", "fc_reg.weight"):
        assert n in hk
    assert head.fc_cls.out_features == 4 and head.fc_reg.out_features == 12
    head.init_weights()
    ext = A.SingleRoIExtractor(dict(type='RoIAlign', out_size=7, sample_num=0), 16, [4, 8, 16, 32, 64])
    assert ext.num_inputs == 5 and ext.roi_layers[0].out_size == (7, 7)
    assert abs(ext.roi_layers[2].spatial_scale - 1 / 16) < 1e-12 and not ext.fp16_enabled
    assert "aligned=True" in repr(ext.roi_layers[0])
    rh = A.StandardRoIHead(
        bbox_roi_extractor=dict(type='SingleRoIExtractor',
                                roi_layer=dict(type='RoIAlign', out_size=7, sample_num=0),
                                out_channels=16, featmap_strides=[4, 8, 16, 32, 64]),
        bbox_head=dict(type='MultiRoIsBBoxHead', in_channels=16, fc_out_channels=32,
                       roi_feat_size=7, num_classes=3))
    assert isinstance(rh.bbox_head, A.MultiRoIsBBoxHead)
