// Exact fp32 geometry shared by the RoI kernels: region boxes, RoI -> level
// map, RoIAlign sampling grid and bilinear taps.
//
// Everything here decides INTEGERS (levels, grid sizes, tap rows/columns) that
// must be bit-identical to the reference evaluated on the CPU, so every
// floating-point step is a single IEEE fp32 operation in the reference's
// order, spelled with __f*_rn intrinsics (which nvcc never contracts to FMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace arfe {

struct RegionBox {
  float b, x1, y1, x2, y2;
};

// Region r of RoI `roi` (r = 0 original, 1 adaptive_w "lw", 2 adaptive_h "lh":
// the channel order of torch.cat([ori, lw, lh]), standard_roi_head.py:155).
// Follows get_adaptive_scale_rois, mmdet/models/utils/additional.py:38-71.
__device__ __forceinline__ RegionBox region_box(const float* __restrict__ roi,
                                                int r, float facs) {
  RegionBox o;
  o.b = roi[0];
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  if (r == 0) {
    o.x1 = x1; o.y1 = y1; o.x2 = x2; o.y2 = y2;
    return o;
  }
  const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f);              // :40
  const float cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);              // :41
  const float rw = __fadd_rn(__fsub_rn(x2, x1), 1.0f);              // :42
  const float rh = __fadd_rn(__fsub_rn(y2, y1), 1.0f);              // :43
  const float h_rate = __fadd_rn(__fmul_rn(__fdiv_rn(rw, rh), facs), 1.0f);  // :50
  const float large_h = __fmul_rn(rh, h_rate);                      // :52
  const float half_lh = __fmul_rn(large_h, 0.5f);
  float half_w;
  if (r == 1) {  // adaptive_w, :62-68 (y extent uses large_h, :66)
    const float w_rate = __fadd_rn(__fmul_rn(__fdiv_rn(rh, rw), facs), 1.0f);  // :51
    half_w = __fmul_rn(__fmul_rn(rw, w_rate), 0.5f);                // :53
  } else {       // adaptive_h, :56-61
    half_w = __fmul_rn(rw, 0.5f);
  }
  o.x1 = fmaxf(__fsub_rn(cx, half_w), 0.1f);
  o.y1 = fmaxf(__fsub_rn(cy, half_lh), 0.1f);
  o.x2 = __fadd_rn(cx, half_w);
  o.y2 = __fadd_rn(cy, half_lh);
  return o;
}

// floor(log2(v)) exactly as torch evaluates it on the CPU, i.e. floor of the
// correctly rounded fp32 log2: just below 2^k the rounded logarithm already
// equals k for the last j_k floats (j = 0 for k <= 2, 1 for k = 3,4, 2 for
// k = 5..8; derivation and exhaustive check in DESIGN.md / tests).  Only
// k in [1, ARFE_MAX_LEVELS-1] can change a clamped level.
__device__ __forceinline__ int floor_log2_rn(float v) {
  const uint32_t bits = __float_as_uint(v);
  const int e = (int)((bits >> 23) & 0xffu) - 127;
  const uint32_t m = bits & 0x7fffffu;
  const int k = e + 1;
  const uint32_t j = (k >= 5) ? 2u : ((k >= 3) ? 1u : 0u);
  return (m + j >= 0x800000u) ? k : e;
}

// SingleRoIExtractor.map_roi_levels, roi_extractors/single_level.py:68-70,92.
// Returns -1 when the scale is NaN (negative area): such a RoI matches no
// level in the reference and its output row stays zero.
__device__ __forceinline__ int map_roi_level(const RegionBox& bx, int L,
                                             float finest_scale) {
  const float area = __fmul_rn(__fsub_rn(bx.x2, bx.x1), __fsub_rn(bx.y2, bx.y1));
  const float scale = __fsqrt_rn(area);
  const float v = __fadd_rn(__fdiv_rn(scale, finest_scale), 1e-6f);
  if (!(v == v)) return -1;
  if (v >= 256.0f) return L - 1;  // also +inf
  if (v < 1.0f) return 0;         // log2 < 0 clamps to level 0
  int l = floor_log2_rn(v);
  return l < 0 ? 0 : (l > L - 1 ? L - 1 : l);
}

// Per-RoI sampling geometry, roi_align_kernel_v2.cu:79-106 (aligned = true).
struct RoiGeom {
  int batch;
  float start_h, start_w, bin_h, bin_w;
  int grid_h, grid_w;
  float count;
};

__device__ __forceinline__ RoiGeom roi_geometry(const RegionBox& bx, float scale,
                                                int PH, int PW,
                                                int sampling_ratio) {
  RoiGeom g;
  g.batch = (int)bx.b;
  g.start_w = __fsub_rn(__fmul_rn(bx.x1, scale), 0.5f);
  g.start_h = __fsub_rn(__fmul_rn(bx.y1, scale), 0.5f);
  const float end_w = __fsub_rn(__fmul_rn(bx.x2, scale), 0.5f);
  const float end_h = __fsub_rn(__fmul_rn(bx.y2, scale), 0.5f);
  const float rw = __fsub_rn(end_w, g.start_w);
  const float rh = __fsub_rn(end_h, g.start_h);
  g.bin_h = __fdiv_rn(rh, (float)PH);
  g.bin_w = __fdiv_rn(rw, (float)PW);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bin_h);
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bin_w);
  // A RoI with negative extent yields an empty grid and a zero output, as in
  // the reference CUDA kernel (the CPU kernel asserts, roi_align_v2.cpp:132).
  if (g.grid_h < 0) g.grid_h = 0;
  if (g.grid_w < 0) g.grid_w = 0;
  const int n = g.grid_h * g.grid_w;
  g.count = (float)(n > 1 ? n : 1);
  return g;
}

// One bilinear sample along one axis (roi_align_kernel_v2.cu:15-50, :111-117):
// neighbouring rows lo/hi and their weights wl (for lo) / wh (for hi);
// lo = -1 when the sample lies outside [-1, extent].
struct AxisTap {
  int lo, hi;
  float wl, wh;
};

__device__ __forceinline__ AxisTap axis_sample(float start, int p, float bin,
                                               int i, int grid, int extent) {
  AxisTap t;
  float pos = __fadd_rn(
      __fadd_rn(start, __fmul_rn((float)p, bin)),
      __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
  if (pos < -1.0f || pos > (float)extent) {
    t.lo = t.hi = -1;
    t.wl = t.wh = 0.f;
    return t;
  }
  if (pos <= 0.f) pos = 0.f;
  int lo = (int)pos;
  int hi;
  if (lo >= extent - 1) {
    hi = lo = extent - 1;
    pos = (float)lo;
  } else {
    hi = lo + 1;
  }
  const float l = __fsub_rn(pos, (float)lo);
  t.lo = lo; t.hi = hi;
  t.wl = __fsub_rn(1.0f, l);
  t.wh = l;
  return t;
}

}  // namespace arfe
