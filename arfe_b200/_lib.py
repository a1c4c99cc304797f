"""ctypes binding of libarfe_b200.so (ABI: include/arfe_b200.h).

There is no CPU or PyTorch fallback: if the library is missing the import of
any op raises, and every non-zero return code becomes a RuntimeError carrying
arfe_last_error() -- the same surface the reference's TORCH_CHECK failures
have (mmdet/ops/roi_align/src/roi_align_ext.cpp:49-54).
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# ARFE_B200_LIB: load another build of the same ABI (scripts/ point it at the
# -DARFE_PROFILE library libarfe_b200_prof.so); the package default has no knobs.
LIB_PATH = os.environ.get("ARFE_B200_LIB") or os.path.join(_PKG, "libarfe_b200.so")

ARFE_F32, ARFE_BF16 = 0, 1
ARFE_NCHW, ARFE_NHWC = 0, 1
MAX_LEVELS = 8
MAX_POOL = 32

c_int, c_float, c_void_p, c_i64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64
_pp = ctypes.POINTER(c_void_p)
_ip = ctypes.POINTER(ctypes.c_int32)
_fp = ctypes.POINTER(c_float)

_SIGNATURES = {
    "arfe_version": ([], c_int),
    "arfe_last_error": ([], ctypes.c_char_p),
    "arfe_roi_fuse_forward": ([_pp, _ip, _ip, _fp, c_int, c_int, c_int, c_void_p, c_int, c_int,
                               c_float, c_int, c_int, c_int, c_float, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "arfe_roi_fuse_backward": ([c_void_p, c_int, _ip, _ip, _fp, c_int, c_int, c_int, c_void_p,
                                c_int, c_int, c_float, c_int, c_int, c_int, c_float, c_int, c_int,
                                _pp, c_void_p], c_int),
    "arfe_roi_fuse_pull_workspace_bytes": ([c_int, c_int, c_int, c_int, _ip, _ip], ctypes.c_size_t),
    "arfe_roi_plan_bytes": ([c_int, c_int, c_int, c_int, _ip, _ip], ctypes.c_size_t),
    "arfe_roi_fuse_forward_plan": ([_pp, _ip, _ip, _fp, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                    c_float, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, ctypes.c_size_t, c_int, c_void_p], c_int),
    "arfe_roi_plan_build": ([_ip, _ip, _fp, c_int, c_int, c_int, c_void_p, c_int, c_int, c_float, c_int,
                             c_int, c_int, c_float, c_int, c_void_p, ctypes.c_size_t, c_void_p], c_int),
    "arfe_roi_pull_bin": ([_ip, _ip, _fp, c_int, c_int, c_int, c_void_p, c_int, c_int, c_float, c_int,
                           c_int, c_int, c_float, c_int, c_int, c_void_p, ctypes.c_size_t, c_void_p], c_int),
    "arfe_roi_fuse_forward_plan_split": ([_pp, _ip, _ip, _fp, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                          c_float, c_int, c_int, c_int, c_float, c_int, _pp,
                                          c_void_p, ctypes.c_size_t, c_int, c_void_p], c_int),
    "arfe_roi_fuse_backward_pull_split": ([_pp, _ip, _ip, _fp, c_int, c_int, c_int, c_void_p,
                                           c_int, c_int, c_float, c_int, c_int, c_int, c_float, c_int,
                                           _pp, c_void_p, ctypes.c_size_t, c_int, c_void_p], c_int),
    "arfe_roi_fuse_backward_pull": ([c_void_p, _ip, _ip, _fp, c_int, c_int, c_int, c_void_p,
                                     c_int, c_int, c_float, c_int, c_int, c_int, c_float, c_int,
                                     _pp, c_void_p, ctypes.c_size_t, c_int, c_void_p], c_int),
    "arfe_roi_align_forward": ([c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p], c_int),
    "arfe_roi_align_backward": ([c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p], c_int),
    "arfe_roi_fuse_taps": ([_ip, _ip, _fp, c_int, c_void_p, c_int, c_int, c_float, c_int, c_int,
                            c_int, c_float, c_int] + [c_void_p] * 11 + [c_void_p], c_int),
    "arfe_rff_gate_forward": ([c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int,
                               c_void_p], c_int),
    "arfe_rff_gate_backward": ([c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64,
                                c_void_p, c_i64, c_i64, c_int, c_void_p], c_int),
    "arfe_rff_softmax_fuse_forward": ([_pp, ctypes.POINTER(c_i64), c_void_p, ctypes.POINTER(c_i64), c_void_p,
                                       ctypes.POINTER(c_i64), c_i64, c_int, c_int, c_int, c_void_p], c_int),
    "arfe_rff_softmax_fuse_backward": ([c_void_p, _pp, ctypes.POINTER(c_i64), c_void_p, ctypes.POINTER(c_i64),
                                        ctypes.POINTER(c_i64), _pp, c_void_p, c_i64, c_int, c_int, c_int,
                                        c_void_p], c_int),
    "arfe_fpn_gate_conv_workspace_bytes": ([c_int, c_int, _ip, _ip], ctypes.c_size_t),
    "arfe_fpn_gate_conv_forward": ([_pp, _pp, _pp, _pp, _pp, _ip, _ip, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    ctypes.c_size_t, _pp, _pp, c_void_p], c_int),
    "arfe_fpn_gate_conv_backward": ([_pp, _pp, _pp, _pp, _pp, _ip, _ip, c_int, c_int, c_int, c_int, c_int, _pp, _pp,
                                     _pp, _pp, _pp, c_void_p], c_int),
    "arfe_nonlocal_backward_rows": ([c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_float, c_void_p], c_int),
    "arfe_nonlocal_default_split": ([c_int, c_int], c_int),
    "arfe_nonlocal_workspace_bytes": ([c_int, c_int, c_int, c_int], ctypes.c_size_t),
    "arfe_nonlocal_attention_forward": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_float, c_int, c_void_p, ctypes.c_size_t, c_void_p], c_int),
    "arfe_nms_workspace_bytes": ([c_int], ctypes.c_size_t),
    "arfe_nms": ([c_void_p, c_int, c_float, c_void_p, ctypes.c_size_t, c_void_p, c_void_p, c_void_p], c_int),
    "arfe_bbox2roi": ([_pp, _ip, _ip, c_int, c_void_p, c_void_p], c_int),
    "arfe_fpn_gather_forward": ([_pp, _ip, _ip, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p], c_int),
    "arfe_fpn_gather_backward": ([c_void_p, c_void_p, _ip, _ip, c_int, c_int, c_int, c_int, c_int,
                                  c_int, _pp, c_void_p], c_int),
    "arfe_fpn_gather_backward_acc": ([c_void_p, c_void_p, _ip, _ip, c_int, c_int, c_int, c_int, c_int,
                                      c_int, _pp, _pp, c_void_p], c_int),
    "arfe_fpn_apply_forward": ([_pp, c_void_p, _pp, _pp, _ip, _ip, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, _pp, c_void_p], c_int),
    "arfe_fpn_apply_backward": ([_pp, c_void_p, _pp, _pp, _ip, _ip, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_void_p, _pp, _pp, c_void_p], c_int),
    "arfe_fpn_backward_fused": ([_pp, c_int, c_void_p, _pp, _pp, c_void_p, c_void_p, _ip, _ip, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_void_p, _pp, _pp, _pp, c_void_p], c_int),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"arfe_b200: {LIB_PATH} is missing. Build it with "
                "`python -m arfe_b200.build` (nvcc, sm_100a). There is no CPU "
                "or PyTorch fallback for these ops.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if a symbol is missing
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().arfe_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def int_array(vals):
    return (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])


def i64_array(vals):
    return (c_i64 * len(vals))(*[int(v) for v in vals])


def float_array(vals):
    return (c_float * len(vals))(*[float(v) for v in vals])


def stream_ptr(device):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(t):
    if t.dtype == torch.float32:
        return ARFE_F32
    if t.dtype == torch.bfloat16:
        return ARFE_BF16
    raise TypeError(f"arfe_b200 supports float32 and bfloat16, got {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "arfe_b200 ops run on CUDA (sm_100a) tensors only; got a "
                f"{t.device} tensor. There is no CPU implementation.")


def layout_of(t):
    """ARFE_NHWC for a channels_last-contiguous 4-d tensor, else ARFE_NCHW
    (the tensor must then be contiguous)."""
    if t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous() and \
            t.is_contiguous(memory_format=torch.channels_last):
        return ARFE_NHWC
    return ARFE_NCHW


def as_layout(t, layout):
    if layout == ARFE_NHWC:
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()
