// Device code shared by the RoI kernels (NCHW compatibility path in
// roi_fuse.cu, channels-last fast path in roi_fuse_nhwc.cu): per-CTA header,
// the separable aggregated tap tables and their construction.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "geometry.cuh"
#include "launch.h"

namespace arfe {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kAxisCap = 512;    // aggregated weights per axis
constexpr int kPitch = 33;       // smem pitch of one staged pixel (32 ch + 1)
constexpr int kChunk = 32;       // channels per staged chunk

struct AxisTable {
  int first[kMaxPool];  // first feature row (column) touched by bin p
  int cnt[kMaxPool];    // number of consecutive rows touched (0: none)
  int off[kMaxPool];    // offset of bin p's weights in w[]
  float w[kAxisCap];
};

struct CtaHeader {
  RoiGeom g;
  int lvl;
  int H, W;
  int ymin, ymax, xmin, xmax;  // union window over all bins (inclusive)
  int overflow;                // tables did not fit -> generic path
  int max_rows;                // max over ph of cnt
};

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Build the aggregated per-axis table with one warp (lane p = bin p).
__device__ inline void build_axis_table(AxisTable& t, int P, float start, float bin,
                                 int grid, int extent, int* overflow,
                                 int lane) {
  int lo_min = 0x7fffffff, hi_max = -1;
  if (lane < P) {
    for (int i = 0; i < grid; ++i) {
      AxisTap s = axis_sample(start, lane, bin, i, grid, extent);
      if (s.lo >= 0) {
        lo_min = min(lo_min, s.lo);
        hi_max = max(hi_max, s.hi);
      }
    }
  }
  int n = (hi_max >= 0) ? (hi_max - lo_min + 1) : 0;
  int incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int off = incl - n;
  if (total > kAxisCap) {
    if (lane == 0) *overflow = 1;
    return;
  }
  if (lane < P) {
    t.first[lane] = (n > 0) ? lo_min : 0;
    t.cnt[lane] = n;
    t.off[lane] = off;
    for (int j = 0; j < n; ++j) t.w[off + j] = 0.f;
    for (int i = 0; i < grid; ++i) {
      AxisTap s = axis_sample(start, lane, bin, i, grid, extent);
      if (s.lo >= 0) {
        t.w[off + s.lo - lo_min] += s.wl;
        t.w[off + s.hi - lo_min] += s.wh;
      }
    }
  }
}

// Header + tables for CTA (k, r).  Returns false when the output row is all
// zeros (no level, batch index out of range, or empty sampling window).
__device__ inline bool setup_cta(const RoiFuseParams& p, int k, int r, CtaHeader& hd,
                          AxisTable& ty, AxisTable& tx) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    RegionBox bx = region_box(p.rois + 5 * (size_t)k, r, p.facs);
    int lvl = (p.L == 1) ? 0 : map_roi_level(bx, p.L, p.finest_scale);
    hd.lvl = lvl;
    hd.overflow = 0;
    if (lvl >= 0) {
      hd.g = roi_geometry(bx, p.scale[lvl], p.PH, p.PW, p.sampling_ratio);
      hd.H = p.H[lvl];
      hd.W = p.W[lvl];
      if (hd.g.batch < 0 || hd.g.batch >= p.B) hd.lvl = -2;
    }
    if (p.lvl_out) p.lvl_out[(size_t)r * p.K + k] = lvl;
    if (p.boxes_out) {
      float* o = p.boxes_out + ((size_t)r * p.K + k) * 5;
      o[0] = bx.b; o[1] = bx.x1; o[2] = bx.y1; o[3] = bx.x2; o[4] = bx.y2;
    }
  }
  __syncthreads();
  if (hd.lvl < 0) return false;
  const int warp = tid >> 5, lane = tid & 31;
  if (warp == 0)
    build_axis_table(ty, p.PH, hd.g.start_h, hd.g.bin_h, hd.g.grid_h, hd.H,
                     &hd.overflow, lane);
  else if (warp == 1)
    build_axis_table(tx, p.PW, hd.g.start_w, hd.g.bin_w, hd.g.grid_w, hd.W,
                     &hd.overflow, lane);
  __syncthreads();
  if (hd.overflow) return true;
  if (tid == 0) {
    int ymin = 0x7fffffff, ymax = -1, xmin = 0x7fffffff, xmax = -1, mr = 0;
    for (int q = 0; q < p.PH; ++q)
      if (ty.cnt[q] > 0) {
        ymin = min(ymin, ty.first[q]);
        ymax = max(ymax, ty.first[q] + ty.cnt[q] - 1);
        mr = max(mr, ty.cnt[q]);
      }
    for (int q = 0; q < p.PW; ++q)
      if (tx.cnt[q] > 0) {
        xmin = min(xmin, tx.first[q]);
        xmax = max(xmax, tx.first[q] + tx.cnt[q] - 1);
      }
    hd.ymin = ymin; hd.ymax = ymax; hd.xmin = xmin; hd.xmax = xmax;
    hd.max_rows = mr;
  }
  __syncthreads();
  return hd.ymax >= 0 && hd.xmax >= 0;
}

// Generic path: reference loop order, direct global taps (any size, slow).
template <typename T, bool kNHWC>
__device__ void forward_generic(const RoiFuseParams& p, const CtaHeader& hd,
                                T* __restrict__ out_blk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int PHW = p.PH * p.PW;
  const T* __restrict__ f = static_cast<const T*>(p.feats[hd.lvl]);
  const int H = hd.H, W = hd.W, C = p.C;
  const RoiGeom& g = hd.g;
  const int nwarps = blockDim.x >> 5;
  for (int bin = warp; bin < PHW; bin += nwarps) {
    const int ph = bin / p.PW, pw = bin % p.PW;
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        AxisTap a = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H);
        if (a.lo < 0) continue;
        for (int ix = 0; ix < g.grid_w; ++ix) {
          AxisTap b = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W);
          if (b.lo < 0) continue;
          size_t i1, i2, i3, i4;
          if (kNHWC) {
            const size_t base = (size_t)g.batch * H * W;
            i1 = (base + (size_t)a.lo * W + b.lo) * C + c;
            i2 = (base + (size_t)a.lo * W + b.hi) * C + c;
            i3 = (base + (size_t)a.hi * W + b.lo) * C + c;
            i4 = (base + (size_t)a.hi * W + b.hi) * C + c;
          } else {
            const size_t base = ((size_t)g.batch * C + c) * H * W;
            i1 = base + (size_t)a.lo * W + b.lo;
            i2 = base + (size_t)a.lo * W + b.hi;
            i3 = base + (size_t)a.hi * W + b.lo;
            i4 = base + (size_t)a.hi * W + b.hi;
          }
          acc += a.wl * b.wl * to_f(f[i1]) + a.wl * b.wh * to_f(f[i2]) +
                 a.wh * b.wl * to_f(f[i3]) + a.wh * b.wh * to_f(f[i4]);
        }
      }
      out_blk[(size_t)c * PHW + bin] = from_f<T>(__fdiv_rn(acc, g.count));
    }
  }
}

// Vector float reductions to global memory (sm_90+): one L2 atomic transaction
// for 4 (2) consecutive floats; the address must be 16 (8) byte aligned.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// px -> (row, col) of a window `ww` pixels wide without an integer division.
__device__ __forceinline__ void split_px(int px, int ww, float inv_ww, int& row,
                                         int& col) {
  row = (int)(((float)px + 0.5f) * inv_ww);
  col = px - row * ww;
  if (col < 0) { --row; col += ww; }
  else if (col >= ww) { ++row; col -= ww; }
}

template <typename T>
__device__ void zero_block(T* __restrict__ out_blk, int n) {
  for (int i = threadIdx.x; i < n; i += kThreads) out_blk[i] = from_f<T>(0.f);
}

// Per-region record written by roi_prep (roi_fuse_nhwc.cu) into the workspace.
struct __align__(16) RegionHdr {
  int lvl;        // level, or -1: contributes nothing
  int batch;
  int ymin, ymax, xmin, xmax;
  int src;        // k * R + r (row of dout)
  int flags;      // bit 0: tables did not fit -> atomic fallback kernel; bits 8-11 / 12-15: bin blocks per row / column
};

constexpr int kHdrBytes = 128 + 2 * (int)sizeof(AxisTable);
static_assert(sizeof(CtaHeader) <= 128, "header must fit its slot");
constexpr int kMaxSmem = 220 * 1024;

template <typename K>
static cudaError_t set_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace arfe
