"""Build recipe for the parity checkers (TEST INFRASTRUCTURE, not product code).

Two artefacts, both git-ignored, both shipped to the GPU box by gpurun:

* ``oracle/liboracle_roialign.so`` -- our plain-C restatement
  (``oracle/roi_align_oracle.c``), gcc, no dependencies.
* ``oracle/_ref/roi_align_ext*.so`` -- the REFERENCE's own RoIAlign, compiled
  unmodified and in place from
  ``/root/reference/mmdet/ops/roi_align/src/roi_align_ext.cpp`` and
  ``.../src/cpu/roi_align_v2.cpp`` (CPU-only: ``WITH_CUDA`` is not defined) as a
  torch C++ extension.  No reference source is copied into this repo; only the
  built ``.so`` lands in ``oracle/_ref/``.  Skipped when ``/root/reference`` is
  absent (the GPU box), where the pre-built file is used.

Run:  python oracle/build_oracle.py
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ARFE_REFERENCE_ROOT", "/root/reference")
REF_SRCS = [
    "mmdet/ops/roi_align/src/roi_align_ext.cpp",
    "mmdet/ops/roi_align/src/cpu/roi_align_v2.cpp",
]
REF_DIR = os.path.join(HERE, "_ref")
C_LIB = os.path.join(HERE, "liboracle_roialign.so")
C_SRC = os.path.join(HERE, "roi_align_oracle.c")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_c_oracle(force=False):
    if not force and _newer(C_LIB, [C_SRC]):
        return C_LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
           "-fno-fast-math", "-fopenmp", "-o", C_LIB, C_SRC, "-lm"]
    subprocess.check_call(cmd)
    return C_LIB


def ref_ext_path():
    hits = sorted(glob.glob(os.path.join(REF_DIR, "roi_align_ext*.so")))
    return hits[0] if hits else None


def build_reference_ext(force=False):
    """Compile the reference's own two C++ files where they lie."""
    srcs = [os.path.join(REF_ROOT, s) for s in REF_SRCS]
    if not all(os.path.exists(s) for s in srcs):
        return ref_ext_path()  # GPU box: use the prebuilt .so if it travelled
    have = ref_ext_path()
    if have and not force and _newer(have, srcs):
        return have
    os.makedirs(REF_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name="roi_align_ext", sources=srcs, build_directory=REF_DIR,
         extra_cflags=["-O2"], verbose=False, is_python_module=True)
    return ref_ext_path()


REF_CUDA_SRC = "mmdet/ops/roi_align/src/cuda/roi_align_kernel_v2.cu"
REF_CUDA_DIR = os.path.join(REF_DIR, "cuda")


def ref_cuda_ext_path():
    hits = sorted(glob.glob(os.path.join(REF_CUDA_DIR, "roi_align_ref_cuda*.so")))
    return hits[0] if hits else None


def build_reference_cuda_ext(force=False):
    """The reference's own CUDA RoIAlign v2 kernel, compiled unmodified and in
    place for sm_100a, bound by oracle/ref_cuda_glue.cpp: the GPU baseline of
    scripts/ref_gpu_baseline.py ('GPU baseline to beat')."""
    src = os.path.join(REF_ROOT, REF_CUDA_SRC)
    glue = os.path.join(HERE, "ref_cuda_glue.cpp")
    if not os.path.exists(src):
        return ref_cuda_ext_path()  # GPU box: use the prebuilt .so if it travelled
    have = ref_cuda_ext_path()
    if have and not force and _newer(have, [src, glue]):
        return have
    os.makedirs(REF_CUDA_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="roi_align_ref_cuda", sources=[glue, src], build_directory=REF_CUDA_DIR,
         extra_cflags=["-O2"], extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"],
         verbose=False, is_python_module=True)
    return ref_cuda_ext_path()


REF_NMS_SRCS = ["mmdet/ops/nms/src/nms_ext.cpp", "mmdet/ops/nms/src/cpu/nms_cpu.cpp"]
REF_NMS_DIR = os.path.join(REF_DIR, "nms")


def ref_nms_ext_path():
    hits = sorted(glob.glob(os.path.join(REF_NMS_DIR, "nms_ext*.so")))
    return hits[0] if hits else None


def build_reference_nms_ext(force=False):
    """The reference's own NMS (nms_ext.cpp + cpu/nms_cpu.cpp, CPU-only: WITH_CUDA is not
    defined), compiled unmodified and in place: the checker of arfe_nms."""
    srcs = [os.path.join(REF_ROOT, s) for s in REF_NMS_SRCS]
    if not all(os.path.exists(s) for s in srcs):
        return ref_nms_ext_path()
    have = ref_nms_ext_path()
    if have and not force and _newer(have, srcs):
        return have
    os.makedirs(REF_NMS_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name="nms_ext", sources=srcs, build_directory=REF_NMS_DIR, extra_cflags=["-O2"], verbose=False,
         is_python_module=True)
    return ref_nms_ext_path()


def main():
    force = "--force" in sys.argv
    print("C oracle      :", build_c_oracle(force))
    print("reference ext :", build_reference_ext(force) or
          "unavailable (no /root/reference and no prebuilt oracle/_ref)")
    print("reference NMS :", build_reference_nms_ext(force) or "unavailable")
    if "--cuda" in sys.argv:
        print("reference CUDA:", build_reference_cuda_ext(force) or "unavailable")


if __name__ == "__main__":
    main()
