"""arfe_b200 -- B200 (sm_100a) implementation of ARFE's region-aware feature
path: AR-FPN aggregation (WFPNDualSpatial) and AR-RFF RoI fusion
(SingleRoIExtractor x 3 regions + MultiRoIsBBoxHead gate), behind the
reference's module/operator surface.  Compute lives in libarfe_b200.so
(C ABI: include/arfe_b200.h); there is no CPU fallback.
"""
from ._compat import (ConvModule, accelerate_class, optimize_detector,  # noqa: F401
                      register_into_mmdet)
from .bbox_head import MultiBBoxHead, MultiRoIsBBoxHead  # noqa: F401
from .functional import (fpn_apply, fpn_gate_conv, fpn_gather, nonlocal_attention, rff_gate,  # noqa: F401
                         roi_fuse, roi_fuse_debug, roi_fuse_split, rff_softmax_fuse, split3)
from .neck import NonLocal2D, WFPNDualSpatial  # noqa: F401
from .proposals import batched_nms, bbox2roi, nms  # noqa: F401
from .regions import get_adaptive_scale_rois  # noqa: F401
from .roi_align import RoIAlign, RoIAlignFunction, roi_align  # noqa: F401
from .roi_extractor import SingleRoIExtractor  # noqa: F401
from .roi_head import CascadeRoIHead, StandardRoIHead  # noqa: F401

__version__ = "0.1.0"
