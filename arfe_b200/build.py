"""Compile arfe_b200/csrc/*.cu for sm_100a.

  libarfe_b200.so       the product (no profiling knobs, no getenv)
  libarfe_b200_prof.so  -DARFE_PROFILE: the same kernels plus the phase-skip / environment
                        knobs the scripts under scripts/ use (never loaded by the package
                        unless ARFE_B200_LIB points at it)

Plain nvcc, no torch headers: the library's ABI is include/arfe_b200.h.  Every
translation unit is compiled to an object in parallel, then linked.
Run:  python -m arfe_b200.build [--force] [--verbose] [--release | --profile]   (default: both)
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libarfe_b200.so")
LIB_PROF = os.path.join(PKG, "libarfe_b200_prof.so")
OBJ = os.path.join(PKG, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(PKG, "..", "include", "arfe_b200.h")]


def _stale(target, inputs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in inputs)


def up_to_date(profile=False):
    return not _stale(LIB_PROF if profile else LIB, sources() + headers())


def build(force=False, verbose=False, profile=False):
    lib = LIB_PROF if profile else LIB
    if not force and up_to_date(profile):
        return lib
    os.makedirs(OBJ, exist_ok=True)
    tag = "prof" if profile else "rel"
    extra = ["-DARFE_PROFILE"] if profile else []
    hdrs = headers()

    def compile_one(src):
        obj = os.path.join(OBJ, f"{os.path.basename(src)[:-3]}.{tag}.o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs)
    return lib


if __name__ == "__main__":
    # both libraries by default (the profile build must never be staler than the product:
    # tests compare the two); --release / --profile build one
    kinds = [False, True]
    if "--release" in sys.argv:
        kinds = [False]
    if "--profile" in sys.argv:
        kinds = [True]
    for prof in kinds:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, profile=prof))
