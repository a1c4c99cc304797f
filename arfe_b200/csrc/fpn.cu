// AR-FPN (WFPNDualSpatial) memory-bound stages.
//   gather : mmdet/models/necks/wfpn_dual_spatial.py:102-113
//            adaptive_max_pool2d (levels below refine) / nearest interpolate
//            (others) to the refine size, summed in level order, divided by L.
//   apply  : mmdet/models/necks/wfpn_dual_spatial.py:118-135
//            out_l = x_l + nearest(bsf -> level size) * (tanh(relu(g1_l)) + tanh(relu(g2_l)))
// The 256->1 3x3 gate convolutions and the NonLocal2D refine stay on PyTorch
// (SURVEY.md section 8(a) rows a2, a3); these kernels consume their outputs.
//
// Index rules (ATen, verified in the survey):
//   adaptive max pool window of output i: [floor(i*in/out), ceil((i+1)*in/out)),
//     first maximum wins, NaN propagates;
//   nearest source of destination d: min(floor(d * float(in)/float(out)), in-1).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "launch.h"

namespace arfe {
namespace {

constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

template <bool kNHWC>
__device__ __forceinline__ size_t at(int b, int c, int y, int x, int C, int H, int W) {
  return kNHWC ? (((size_t)b * H + y) * W + x) * C + c
               : (((size_t)b * C + c) * H + y) * W + x;
}

// Decode a flat index in memory order of a [B,C,H,W] tensor.
template <bool kNHWC>
__device__ __forceinline__ void decode(size_t i, int C, int H, int W, int& b, int& c,
                                       int& y, int& x) {
  if (kNHWC) {
    c = (int)(i % C); i /= C;
    x = (int)(i % W); i /= W;
    y = (int)(i % H); b = (int)(i / H);
  } else {
    x = (int)(i % W); i /= W;
    y = (int)(i % H); i /= H;
    c = (int)(i % C); b = (int)(i / C);
  }
}

__device__ __forceinline__ int pool_start(int i, int in, int out) {
  return (int)(((long long)i * in) / out);
}
__device__ __forceinline__ int pool_end(int i, int in, int out) {
  return (int)(((long long)(i + 1) * in + out - 1) / out);
}
__device__ __forceinline__ int nearest_src(int d, int in, int out) {
  const float scale = __fdiv_rn((float)in, (float)out);
  const int s = (int)floorf(__fmul_rn((float)d, scale));
  return s < in - 1 ? s : in - 1;
}

// ---------------------------------------------------------------- gather fwd
template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads)
gather_fwd(const FpnParams p) {
  const int Hr = p.Hr, Wr = p.Wr, C = p.C;
  const size_t total = (size_t)p.B * C * Hr * Wr;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  int b, c, Y, X;
  decode<kNHWC>(i, C, Hr, Wr, b, c, Y, X);
  float acc = 0.f;
  for (int l = 0; l < p.L; ++l) {
    const T* __restrict__ f = static_cast<const T*>(p.feats[l]);
    const int H = p.H[l], W = p.W[l];
    float v;
    if (l < p.refine_level) {
      const int y0 = pool_start(Y, H, Hr), y1 = pool_end(Y, H, Hr);
      const int x0 = pool_start(X, W, Wr), x1 = pool_end(X, W, Wr);
      float best = -CUDART_INF_F;
      int arg = 0;
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const float t = ldf(f + at<kNHWC>(b, c, y, x, C, H, W));
          if (t > best || t != t) { best = t; arg = (y - y0) * (x1 - x0) + (x - x0); }
        }
      v = best;
      if (p.argmax)
        p.argmax[(((size_t)l * p.B + b) * C + c) * Hr * Wr + (size_t)Y * Wr + X] = (uint8_t)arg;
    } else {
      v = ldf(f + at<kNHWC>(b, c, nearest_src(Y, H, Hr), nearest_src(X, W, Wr), C, H, W));
    }
    acc = __fadd_rn(acc, v);
  }
  stf(static_cast<T*>(p.gathered) + i, __fdiv_rn(acc, (float)p.L));
}

// ---------------------------------------------------------------- gather bwd
// One thread per element of every level's gradient (all fully written).
struct LevelOffsets { size_t start[kMaxLevels + 1]; };

template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads)
gather_bwd(const FpnParams p, const LevelOffsets lo) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lo.start[p.L]) return;
  int l = 0;
  while (l + 1 < p.L && i >= lo.start[l + 1]) ++l;
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  int b, c, y, x;
  decode<kNHWC>(i - lo.start[l], C, H, W, b, c, y, x);
  const T* __restrict__ d = static_cast<const T*>(p.gathered);  // dout [B,C,Hr,Wr]
  float g = 0.f;
  if (l < p.refine_level) {
    // output cells whose pooling window contains (y, x)
    int Y0 = (int)(((long long)y * Hr) / H);
    while (Y0 > 0 && pool_end(Y0 - 1, H, Hr) > y) --Y0;
    int X0 = (int)(((long long)x * Wr) / W);
    while (X0 > 0 && pool_end(X0 - 1, W, Wr) > x) --X0;
    for (int Y = Y0; Y < Hr && pool_start(Y, H, Hr) <= y; ++Y) {
      const int y0 = pool_start(Y, H, Hr);
      if (pool_end(Y, H, Hr) <= y) continue;
      for (int X = X0; X < Wr && pool_start(X, W, Wr) <= x; ++X) {
        const int x0 = pool_start(X, W, Wr), x1 = pool_end(X, W, Wr);
        if (x1 <= x) continue;
        const int arg = p.argmax[(((size_t)l * p.B + b) * C + c) * Hr * Wr + (size_t)Y * Wr + X];
        if (arg == (y - y0) * (x1 - x0) + (x - x0))
          g += ldf(d + at<kNHWC>(b, c, Y, X, C, Hr, Wr));
      }
    }
  } else {
    // destinations (Y, X) whose nearest source is (y, x): contiguous ranges
    const float sy = (float)Hr / (float)H, sx = (float)Wr / (float)W;
    int Ya = (int)floorf((float)y * sy) - 1; if (Ya < 0) Ya = 0;
    int Yb = (int)ceilf((float)(y + 1) * sy) + 1; if (Yb > Hr) Yb = Hr;
    int Xa = (int)floorf((float)x * sx) - 1; if (Xa < 0) Xa = 0;
    int Xb = (int)ceilf((float)(x + 1) * sx) + 1; if (Xb > Wr) Xb = Wr;
    for (int Y = Ya; Y < Yb; ++Y) {
      if (nearest_src(Y, H, Hr) != y) continue;
      for (int X = Xa; X < Xb; ++X)
        if (nearest_src(X, W, Wr) == x) g += ldf(d + at<kNHWC>(b, c, Y, X, C, Hr, Wr));
    }
  }
  stf(static_cast<T*>(p.outs[l]) + (i - lo.start[l]), __fdiv_rn(g, (float)p.L));
}

// ----------------------------------------------------------------- apply fwd
__device__ __forceinline__ float gate_value(float a, float b) {
  return tanhf(fmaxf(a, 0.f)) + tanhf(fmaxf(b, 0.f));
}

// NCHW: thread = one pixel of one level, loops over a group of channels so
// the two tanh are amortised; consecutive threads = consecutive x (coalesced).
template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_fwd_nchw(const FpnParams p, const LevelOffsets lo /* pixel offsets */, int cg) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lo.start[p.L]) return;
  int l = 0;
  while (l + 1 < p.L && i >= lo.start[l + 1]) ++l;
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  size_t r = i - lo.start[l];
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H);
  const int b = (int)(r / H);
  const size_t pix = ((size_t)b * H + y) * W + x;
  const float gate = gate_value(ldf(static_cast<const T*>(p.g1[l]) + pix),
                                ldf(static_cast<const T*>(p.g2[l]) + pix));
  const int ny = nearest_src(y, Hr, H), nx = nearest_src(x, Wr, W);
  const int c0 = blockIdx.y * cg, c1 = min(C, c0 + cg);
  const T* __restrict__ xin = static_cast<const T*>(p.feats[l]) + (((size_t)b * C + c0) * H + y) * W + x;
  T* __restrict__ out = static_cast<T*>(p.outs[l]) + (((size_t)b * C + c0) * H + y) * W + x;
  const T* __restrict__ bs = static_cast<const T*>(p.bsf) + (((size_t)b * C + c0) * Hr + ny) * Wr + nx;
  const size_t hw = (size_t)H * W, hwr = (size_t)Hr * Wr;
#pragma unroll 4
  for (int c = 0; c < c1 - c0; ++c)
    stf(out + c * hw, fmaf(ldf(bs + c * hwr), gate, ldf(xin + c * hw)));
}

// NHWC: thread = (pixel, channel) with channel fastest.
template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_fwd_nhwc(const FpnParams p, const LevelOffsets lo /* element offsets */) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lo.start[p.L]) return;
  int l = 0;
  while (l + 1 < p.L && i >= lo.start[l + 1]) ++l;
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  int b, c, y, x;
  decode<true>(i - lo.start[l], C, H, W, b, c, y, x);
  const size_t pix = ((size_t)b * H + y) * W + x;
  const float gate = gate_value(ldf(static_cast<const T*>(p.g1[l]) + pix),
                                ldf(static_cast<const T*>(p.g2[l]) + pix));
  const int ny = nearest_src(y, Hr, H), nx = nearest_src(x, Wr, W);
  const float bv = ldf(static_cast<const T*>(p.bsf) + at<true>(b, c, ny, nx, C, Hr, Wr));
  const size_t e = i - lo.start[l];
  stf(static_cast<T*>(p.outs[l]) + e, fmaf(bv, gate, ldf(static_cast<const T*>(p.feats[l]) + e)));
}

// ----------------------------------------------------------------- apply bwd
// Kernel 1 (per level pixel): d(gate sum) = sum_c dout * bsf -> dg1, dg2.
// Kernel 2 (per refine element): dbsf = sum over levels and over the level
// pixels that read this bsf element of dout * gate.  No atomics, every output
// written exactly once.
template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads)
apply_bwd_gates(const FpnParams p, const LevelOffsets lo /* pixel offsets */) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lo.start[p.L]) return;
  int l = 0;
  while (l + 1 < p.L && i >= lo.start[l + 1]) ++l;
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  size_t r = i - lo.start[l];
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H);
  const int b = (int)(r / H);
  const int ny = nearest_src(y, Hr, H), nx = nearest_src(x, Wr, W);
  const T* __restrict__ d = static_cast<const T*>(p.feats[l]);
  const T* __restrict__ bs = static_cast<const T*>(p.bsf);
  float s = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c)
    s = fmaf(ldf(d + at<kNHWC>(b, c, y, x, C, H, W)),
             ldf(bs + at<kNHWC>(b, c, ny, nx, C, Hr, Wr)), s);
  const size_t pix = ((size_t)b * H + y) * W + x;
  const float a1 = ldf(static_cast<const T*>(p.g1[l]) + pix);
  const float a2 = ldf(static_cast<const T*>(p.g2[l]) + pix);
  const float t1 = tanhf(fmaxf(a1, 0.f)), t2 = tanhf(fmaxf(a2, 0.f));
  p.dg1[l][pix] = a1 > 0.f ? s * (1.f - t1 * t1) : 0.f;
  p.dg2[l][pix] = a2 > 0.f ? s * (1.f - t2 * t2) : 0.f;
}

template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads)
apply_bwd_bsf(const FpnParams p) {
  const int Hr = p.Hr, Wr = p.Wr, C = p.C;
  const size_t total = (size_t)p.B * C * Hr * Wr;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  int b, c, Y, X;
  decode<kNHWC>(i, C, Hr, Wr, b, c, Y, X);
  float acc = 0.f;
  for (int l = 0; l < p.L; ++l) {
    const int H = p.H[l], W = p.W[l];
    // level pixels (y, x) with nearest_src(y: Hr <- H) == Y: contiguous range
    const float sy = (float)H / (float)Hr, sx = (float)W / (float)Wr;
    int ya = (int)floorf((float)Y * sy) - 1; if (ya < 0) ya = 0;
    int yb = (int)ceilf((float)(Y + 1) * sy) + 1; if (yb > H) yb = H;
    int xa = (int)floorf((float)X * sx) - 1; if (xa < 0) xa = 0;
    int xb = (int)ceilf((float)(X + 1) * sx) + 1; if (xb > W) xb = W;
    const T* __restrict__ d = static_cast<const T*>(p.feats[l]);
    const T* __restrict__ q1 = static_cast<const T*>(p.g1[l]);
    const T* __restrict__ q2 = static_cast<const T*>(p.g2[l]);
    for (int y = ya; y < yb; ++y) {
      if (nearest_src(y, Hr, H) != Y) continue;
      for (int x = xa; x < xb; ++x) {
        if (nearest_src(x, Wr, W) != X) continue;
        const size_t pix = ((size_t)b * H + y) * W + x;
        acc = fmaf(ldf(d + at<kNHWC>(b, c, y, x, C, H, W)),
                   gate_value(ldf(q1 + pix), ldf(q2 + pix)), acc);
      }
    }
  }
  p.dbsf[i] = acc;
}

inline LevelOffsets offsets(const FpnParams& p, bool per_pixel) {
  LevelOffsets lo;
  size_t s = 0;
  for (int l = 0; l < p.L; ++l) {
    lo.start[l] = s;
    s += (size_t)p.B * (per_pixel ? 1 : p.C) * p.H[l] * p.W[l];
  }
  for (int l = p.L; l <= kMaxLevels; ++l) lo.start[l] = s;
  return lo;
}
inline unsigned blocks_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

#define ARFE_DISPATCH(KERNEL, GRID, ...)                                                 \
  do {                                                                                   \
    if (dtype == 0) {                                                                    \
      if (layout == 0) KERNEL<float, false><<<GRID, kThreads, 0, stream>>>(__VA_ARGS__); \
      else KERNEL<float, true><<<GRID, kThreads, 0, stream>>>(__VA_ARGS__);              \
    } else {                                                                             \
      if (layout == 0) KERNEL<__nv_bfloat16, false><<<GRID, kThreads, 0, stream>>>(__VA_ARGS__); \
      else KERNEL<__nv_bfloat16, true><<<GRID, kThreads, 0, stream>>>(__VA_ARGS__);      \
    }                                                                                    \
  } while (0)

cudaError_t launch_fpn_gather_forward(const FpnParams& p, int dtype, int layout,
                                      cudaStream_t stream) {
  const size_t total = (size_t)p.B * p.C * p.Hr * p.Wr;
  if (total == 0) return cudaSuccess;
  ARFE_DISPATCH(gather_fwd, blocks_for(total), p);
  return cudaGetLastError();
}

cudaError_t launch_fpn_gather_backward(const FpnParams& p, int dtype, int layout,
                                       cudaStream_t stream) {
  const LevelOffsets lo = offsets(p, false);
  if (lo.start[p.L] == 0) return cudaSuccess;
  ARFE_DISPATCH(gather_bwd, blocks_for(lo.start[p.L]), p, lo);
  return cudaGetLastError();
}

cudaError_t launch_fpn_apply_forward(const FpnParams& p, int dtype, int layout,
                                     cudaStream_t stream) {
  if (layout == 0) {
    const LevelOffsets lo = offsets(p, true);
    if (lo.start[p.L] == 0) return cudaSuccess;
    const int cg = p.C >= 32 ? 32 : p.C;
    dim3 grid(blocks_for(lo.start[p.L]), (p.C + cg - 1) / cg);
    if (dtype == 0) apply_fwd_nchw<float><<<grid, kThreads, 0, stream>>>(p, lo, cg);
    else apply_fwd_nchw<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(p, lo, cg);
  } else {
    const LevelOffsets lo = offsets(p, false);
    if (lo.start[p.L] == 0) return cudaSuccess;
    if (dtype == 0) apply_fwd_nhwc<float><<<blocks_for(lo.start[p.L]), kThreads, 0, stream>>>(p, lo);
    else apply_fwd_nhwc<__nv_bfloat16><<<blocks_for(lo.start[p.L]), kThreads, 0, stream>>>(p, lo);
  }
  return cudaGetLastError();
}

cudaError_t launch_fpn_apply_backward(const FpnParams& p, int dtype, int layout,
                                      cudaStream_t stream) {
  const LevelOffsets lo = offsets(p, true);
  if (lo.start[p.L] == 0) return cudaSuccess;
  ARFE_DISPATCH(apply_bwd_gates, blocks_for(lo.start[p.L]), p, lo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t total = (size_t)p.B * p.C * p.Hr * p.Wr;
  ARFE_DISPATCH(apply_bwd_bsf, blocks_for(total), p);
  return cudaGetLastError();
}

}  // namespace arfe
