// AR-FPN gate convolutions: the two 256 -> 1 3x3 convolutions
// whose outputs gate the residual of every level,
//   g1_l = reduce_convs[l].conv(x_l),  g2_l = reduce_convs2[l].conv(x_l)
//   (mmdet/models/necks/wfpn_dual_spatial.py:38-55, :120-121; bias included, no activation --
//    relu and tanh are fused into the apply kernel).
// A C -> 1 convolution is a channel REDUCTION, bound by reading x, not a contraction worth a
// tensor core: through cuDNN the pyramid is read twice more (once per convolution).  Here x is
// read ONCE for both: phase 1 turns every pixel's C-vector into its 2 x 9 dot products with the
// 18 weight vectors (warp == pixels, lanes == channels, the 18 sums of a pixel reduced together
// with a handful of shuffles), phase 2 adds, per output pixel, the 9 dot products its
// neighbours computed for the tap that points at it (zero padding), plus the bias.
// Channels-last only; all levels in one launch per phase.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.h"

namespace arfe {
namespace {

constexpr int kThreads = 256;
constexpr int kPixPerCta = 256;  // pixels a CTA walks with one copy of the level's weights in shared memory
constexpr int kDots = 18;        // 2 filters x 9 taps

template <typename T, int V>
__device__ __forceinline__ void ld_chan(const T* __restrict__ p, float (&f)[V]) {
  if constexpr (sizeof(T) == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
}

// Warp totals of 16 per-lane values with 15 + 1 shuffles: lane L ends up with the total of
// value (L >> 1) & 15 ... returned for index `want` via one more shuffle by the caller.
__device__ __forceinline__ float reduce16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2], d;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (b4 ? v[8 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[8 + i], 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = (b3 ? a[4 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, b3 ? a[i] : a[4 + i], 8);
#pragma unroll
  for (int i = 0; i < 2; ++i) c[i] = (b2 ? b[2 + i] : b[i]) + __shfl_xor_sync(0xffffffffu, b2 ? b[i] : b[2 + i], 4);
  d = (b1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b1 ? c[0] : c[1], 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;  // lane L holds value index 8*b4 + 4*b3 + 2*b2 + b1 == (L >> 1) & 15
}

struct ConvLevels {
  const void* x[kMaxLevels];
  const float* w1[kMaxLevels];
  const float* w2[kMaxLevels];
  const float* b1[kMaxLevels];
  const float* b2[kMaxLevels];
  void* g1[kMaxLevels];
  void* g2[kMaxLevels];
  int H[kMaxLevels], W[kMaxLevels];
  long long pix0[kMaxLevels + 1];   // first pixel (over B * H * W) of level l in the dots buffer
  int cta0[kMaxLevels + 1];         // first CTA of level l (phase 1)
  int L, B, C;
};

// phase 1: dots[pix][18]
template <typename T, int NV>
__global__ void __launch_bounds__(kThreads)
gate_conv_dots(const ConvLevels lv, float* __restrict__ dots) {
  constexpr int V = sizeof(T) == 4 ? 4 : 8;
  extern __shared__ float wsm[];  // [18][C]
  int l = 0;
  while (l + 1 < lv.L && (int)blockIdx.x >= lv.cta0[l + 1]) ++l;
  const int C = lv.C;
  // the level's 18 weight vectors, transposed from the convolution layout [1][C][3][3]
  for (int e = threadIdx.x; e < kDots * C; e += kThreads) {
    const int d = e / C, c = e - d * C;
    const float* w = d < 9 ? lv.w1[l] : lv.w2[l];
    wsm[e] = __ldg(w + c * 9 + (d < 9 ? d : d - 9));
  }
  __syncthreads();
  const long long npix = (long long)lv.B * lv.H[l] * lv.W[l];
  const long long p0 = (long long)((int)blockIdx.x - lv.cta0[l]) * kPixPerCta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* __restrict__ x = static_cast<const T*>(lv.x[l]);
  float* __restrict__ out = dots + lv.pix0[l] * kDots;
  for (long long pp = p0 + warp; pp < min(p0 + kPixPerCta, npix); pp += kThreads / 32) {
    float xv[NV][V];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * V;
      if (c < C) ld_chan<T, V>(x + pp * C + c, xv[v]);
      else {
#pragma unroll
        for (int u = 0; u < V; ++u) xv[v][u] = 0.f;
      }
    }
    float s[kDots];
#pragma unroll
    for (int d = 0; d < kDots; ++d) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = (v * 32 + lane) * V;
        if (c < C) {
#pragma unroll
          for (int u = 0; u < V; u += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wsm + d * C + c + u);
            a0 = fmaf(xv[v][u], w.x, a0); a1 = fmaf(xv[v][u + 1], w.y, a1);
            a0 = fmaf(xv[v][u + 2], w.z, a0); a1 = fmaf(xv[v][u + 3], w.w, a1);
          }
        }
      }
      s[d] = a0 + a1;
    }
    float first[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) first[d] = s[d];
    const float r = reduce16(first, lane);
    float t16 = s[16], t17 = s[17];
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
      t16 += __shfl_xor_sync(0xffffffffu, t16, dd);
      t17 += __shfl_xor_sync(0xffffffffu, t17, dd);
    }
    // lanes 0, 2, 4 ... 30 hold dots 0 .. 15; lanes 1 and 3 write dots 16 and 17
    float* o = out + pp * kDots;
    if (!(lane & 1)) o[lane >> 1] = r;
    else if (lane == 1) o[16] = t16;
    else if (lane == 3) o[17] = t17;
  }
}

// phase 2: g_f[b, y, x] = bias_f + sum over the 3x3 taps (ky, kx) of dots[b, y+ky-1, x+kx-1][f*9 + ky*3 + kx]
template <typename T>
__global__ void __launch_bounds__(kThreads)
gate_conv_sum(const ConvLevels lv, const float* __restrict__ dots) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lv.pix0[lv.L]) return;
  int l = 0;
  while (l + 1 < lv.L && i >= lv.pix0[l + 1]) ++l;
  const int H = lv.H[l], W = lv.W[l];
  const long long r = i - lv.pix0[l];
  const int x = (int)(r % W), y = (int)((r / W) % H);
  const long long img0 = r - ((long long)y * W + x);  // first pixel of this image
  const float* __restrict__ d = dots + lv.pix0[l] * kDots;
  float s1 = __ldg(lv.b1[l]), s2 = __ldg(lv.b2[l]);
  // the reference sums in cuDNN's order; this order is fixed (taps row-major, filter 1 then 2)
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const float* q = d + (img0 + (long long)yy * W + xx) * kDots + ky * 3 + kx;
      s1 += __ldg(q);
      s2 += __ldg(q + 9);
    }
  }
  if constexpr (sizeof(T) == 4) {
    static_cast<float*>(lv.g1[l])[r] = s1;
    static_cast<float*>(lv.g2[l])[r] = s2;
  } else {
    static_cast<__nv_bfloat16*>(lv.g1[l])[r] = __float2bfloat16_rn(s1);
    static_cast<__nv_bfloat16*>(lv.g2[l])[r] = __float2bfloat16_rn(s2);
  }
}


// ------------------------------------------------------------------------------------------------
// Backward of both gate convolutions of every level in ONE pass over x (r2):
//   g_f[q'] = b_f + sum_{tap, c} w_f[c][tap] x[q' + off(tap)][c]
//   => pixel q of x meets, for filter f and tap t, the output gradient s_{f,t} = dg_f[q - off(t)]
//      (zero outside the map):   dx[q][c]       = sum_{f,t} s_{f,t} w_f[c][t]
//                                dw_f[c][t]    += sum_q     s_{f,t} x[q][c]        db_f += sum_q dg_f[q]
// The 18 scalars of a pixel serve both sums: x is read once, dx written once (through two library
// calls per convolution x is read twice and dx written and summed twice).  Warp == (pixel, group of
// 128 channels), lane == 4 channels: 72 weight-gradient accumulators per lane, reduced over the
// CTA in shared memory and added to dw with one atomic per element and CTA (the order of these
// additions is not fixed: weight gradients are reproducible to rounding, like the library's).
constexpr int kBwdPixPerCta = 512;

struct ConvBwdLevels {
  const void* x[kMaxLevels];
  const float* w1[kMaxLevels];
  const float* w2[kMaxLevels];
  const void* dg1[kMaxLevels];
  const void* dg2[kMaxLevels];
  void* dx[kMaxLevels];
  float* dw1[kMaxLevels];
  float* dw2[kMaxLevels];
  float* db1[kMaxLevels];
  float* db2[kMaxLevels];
  int H[kMaxLevels], W[kMaxLevels];
  int cta0[kMaxLevels + 1];
  int L, B, C, need_dx;
};

template <typename T>
__device__ __forceinline__ float ld_scalar(const T* p) {
  if constexpr (sizeof(T) == 4) return __ldg(p);
  else return __bfloat162float(*p);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
gate_conv_bwd(const ConvBwdLevels lv) {
  extern __shared__ float bsm[];  // [18][C] weights | [18][C] weight-gradient partials | [2] bias partials
  int l = 0;
  while (l + 1 < lv.L && (int)blockIdx.x >= lv.cta0[l + 1]) ++l;
  const int C = lv.C, H = lv.H[l], W = lv.W[l];
  float* wsm = bsm;
  float* dsm = bsm + kDots * C;
  float* bsum = dsm + kDots * C;
  for (int e = threadIdx.x; e < 9 * C; e += kThreads) {  // linear reads of [C][9], transposed in shared memory
    const int cc = e / 9, t = e - cc * 9;
    wsm[t * C + cc] = __ldg(lv.w1[l] + e);
    wsm[(9 + t) * C + cc] = __ldg(lv.w2[l] + e);
  }
  for (int e = threadIdx.x; e < kDots * C; e += kThreads) dsm[e] = 0.f;
  if (threadIdx.x < 2) bsum[threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (C + 127) / 128;            // 1, 2 or 4 channel groups of 128
  const int group = warp % ngroups, slot = warp / ngroups, nslots = (kThreads / 32) / ngroups;
  const int c = group * 128 + lane * 4;
  const bool cact = c < C;
  const long long npix = (long long)lv.B * H * W;
  const long long p0 = (long long)((int)blockIdx.x - lv.cta0[l]) * kBwdPixPerCta;
  const long long p1 = min(p0 + kBwdPixPerCta, npix);
  const T* __restrict__ x = static_cast<const T*>(lv.x[l]);
  const T* __restrict__ g1 = static_cast<const T*>(lv.dg1[l]);
  const T* __restrict__ g2 = static_cast<const T*>(lv.dg2[l]);
  T* __restrict__ dx = static_cast<T*>(lv.dx[l]);
  float acc[kDots][4];
#pragma unroll
  for (int d = 0; d < kDots; ++d)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[d][u] = 0.f;
  float bacc = 0.f;  // lanes 4 and 13 of group 0: the centre taps are dg_f[q] itself
  // one pixel's operands: lane d < 18 fetches s_d = dg_f[q - off(tap)], every lane its 4 channels of x[q]
  auto fetch = [&](long long pp, float& sv, float (&xv)[4]) {
    sv = 0.f;
    xv[0] = xv[1] = xv[2] = xv[3] = 0.f;
    const unsigned pi = (unsigned)pp;  // B * H * W < 2^31 (checked by the launcher): 32-bit divisions
    const int xq = (int)(pi % (unsigned)W), yq = (int)((pi / (unsigned)W) % (unsigned)H);
    if (lane < kDots) {
      const int t = lane < 9 ? lane : lane - 9;
      const int yy = yq - (t / 3 - 1), xx = xq - (t % 3 - 1);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        sv = ld_scalar<T>((lane < 9 ? g1 : g2) + pp + (long long)(yy - yq) * W + (xx - xq));
    }
    if (cact) {
      if constexpr (sizeof(T) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + pp * C + c));
        xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
      } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(x + pp * C + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
        const float2 a = __bfloat1622float2(h[0]), b2 = __bfloat1622float2(h[1]);
        xv[0] = a.x; xv[1] = a.y; xv[2] = b2.x; xv[3] = b2.y;
      }
    }
  };
  // (Keeping the operands of the next two pixels in flight was measured slower, 500 vs 392 us: the
  // kernel is bound by its ~300 instructions per pixel and warp at 16 resident warps, not by the loads.)
  for (long long pp = p0 + slot; pp < p1; pp += nslots) {
    float sv, xv[4];
    fetch(pp, sv, xv);
    if (group == 0 && (lane == 4 || lane == 13)) bacc += sv;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int d = 0; d < kDots; ++d) {
      const float s = __shfl_sync(0xffffffffu, sv, d);
      if (cact) {
        const float4 w = *reinterpret_cast<const float4*>(wsm + d * C + c);
        d0 = fmaf(s, w.x, d0); d1 = fmaf(s, w.y, d1); d2 = fmaf(s, w.z, d2); d3 = fmaf(s, w.w, d3);
      }
      acc[d][0] = fmaf(s, xv[0], acc[d][0]); acc[d][1] = fmaf(s, xv[1], acc[d][1]);
      acc[d][2] = fmaf(s, xv[2], acc[d][2]); acc[d][3] = fmaf(s, xv[3], acc[d][3]);
    }
    if (lv.need_dx && cact) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dx + pp * C + c) = make_float4(d0, d1, d2, d3);
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(d0, d1), hi = __floats2bfloat162_rn(d2, d3);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dx + pp * C + c) = o;
      }
    }
  }
  // CTA-level reduction of the weight gradients, then one atomic per element
  if (cact) {
#pragma unroll
    for (int d = 0; d < kDots; ++d)
#pragma unroll
      for (int u = 0; u < 4; ++u) atomicAdd(dsm + d * C + c + u, acc[d][u]);
  }
  if (group == 0 && (lane == 4 || lane == 13)) atomicAdd(bsum + (lane == 13), bacc);
  __syncthreads();
  for (int e = threadIdx.x; e < kDots * C; e += kThreads) {
    const int d = e / C, cc = e - d * C;
    float* dw = d < 9 ? lv.dw1[l] : lv.dw2[l];
    atomicAdd(dw + cc * 9 + (d < 9 ? d : d - 9), dsm[e]);
  }
  if (threadIdx.x == 0) atomicAdd(lv.db1[l], bsum[0]);
  if (threadIdx.x == 1) atomicAdd(lv.db2[l], bsum[1]);
}

}  // namespace

size_t fpn_gate_conv_workspace_bytes(int L, int B, const int* H, const int* W) {
  size_t pix = 0;
  for (int l = 0; l < L; ++l) pix += (size_t)B * H[l] * W[l];
  return pix * kDots * sizeof(float);
}

cudaError_t launch_fpn_gate_conv_forward(const void* const* feats, const float* const* w1, const float* const* b1,
                                         const float* const* w2, const float* const* b2, const int* H, const int* W,
                                         int L, int B, int C, int dtype, void* workspace, void* const* g1,
                                         void* const* g2, cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const int nv = (C + 32 * V - 1) / (32 * V);
  if (C % V || nv > 2 || (size_t)kDots * C * 4 > 200 * 1024) return cudaErrorNotSupported;
  ConvLevels lv;
  long long pix = 0;
  int cta = 0;
  for (int l = 0; l < kMaxLevels; ++l) {
    lv.pix0[l] = pix;
    lv.cta0[l] = cta;
    if (l < L) {
      lv.x[l] = feats[l]; lv.w1[l] = w1[l]; lv.w2[l] = w2[l]; lv.b1[l] = b1[l]; lv.b2[l] = b2[l];
      lv.g1[l] = g1[l]; lv.g2[l] = g2[l]; lv.H[l] = H[l]; lv.W[l] = W[l];
      const long long n = (long long)B * H[l] * W[l];
      pix += n;
      cta += (int)((n + kPixPerCta - 1) / kPixPerCta);
    } else {
      lv.x[l] = nullptr; lv.w1[l] = lv.w2[l] = lv.b1[l] = lv.b2[l] = nullptr; lv.g1[l] = lv.g2[l] = nullptr;
      lv.H[l] = lv.W[l] = 1;
    }
  }
  lv.pix0[kMaxLevels] = pix; lv.cta0[kMaxLevels] = cta;
  for (int l = L; l <= kMaxLevels; ++l) { lv.pix0[l] = pix; lv.cta0[l] = cta; }
  lv.L = L; lv.B = B; lv.C = C;
  if (pix == 0) return cudaSuccess;
  float* dots = static_cast<float*>(workspace);
  const int smem = kDots * C * 4;
  cudaError_t e;
#define ARFE_DOTS(TT, NV)                                                                                         \
  do {                                                                                                            \
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(gate_conv_dots<TT, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e; \
    gate_conv_dots<TT, NV><<<cta, kThreads, smem, stream>>>(lv, dots);                                            \
  } while (0)
  if (dtype == 0) { if (nv == 1) ARFE_DOTS(float, 1); else ARFE_DOTS(float, 2); }
  else { if (nv == 1) ARFE_DOTS(__nv_bfloat16, 1); else ARFE_DOTS(__nv_bfloat16, 2); }
#undef ARFE_DOTS
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const unsigned grid = (unsigned)((pix + kThreads - 1) / kThreads);
  if (dtype == 0) gate_conv_sum<float><<<grid, kThreads, 0, stream>>>(lv, dots);
  else gate_conv_sum<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(lv, dots);
  return cudaGetLastError();
}

cudaError_t launch_fpn_gate_conv_backward(const void* const* feats, const float* const* w1, const float* const* w2,
                                          const void* const* dg1, const void* const* dg2, const int* H, const int* W,
                                          int L, int B, int C, int dtype, void* const* dx, float* const* dw1,
                                          float* const* db1, float* const* dw2, float* const* db2,
                                          cudaStream_t stream) {
  const size_t smem = ((size_t)2 * kDots * C + 2) * sizeof(float);
  if (C % 4 || C > 512 || smem > 200 * 1024) return cudaErrorNotSupported;
  ConvBwdLevels lv;
  int cta = 0;
  for (int l = 0; l < kMaxLevels; ++l) {
    lv.cta0[l] = cta;
    if (l < L) {
      lv.x[l] = feats[l]; lv.w1[l] = w1[l]; lv.w2[l] = w2[l]; lv.dg1[l] = dg1[l]; lv.dg2[l] = dg2[l];
      lv.dx[l] = dx ? dx[l] : nullptr; lv.dw1[l] = dw1[l]; lv.dw2[l] = dw2[l]; lv.db1[l] = db1[l]; lv.db2[l] = db2[l];
      lv.H[l] = H[l]; lv.W[l] = W[l];
      const long long n = (long long)B * H[l] * W[l];
      cta += (int)((n + kBwdPixPerCta - 1) / kBwdPixPerCta);
    } else {
      lv.x[l] = lv.dg1[l] = lv.dg2[l] = nullptr; lv.w1[l] = lv.w2[l] = nullptr; lv.dx[l] = nullptr;
      lv.dw1[l] = lv.dw2[l] = lv.db1[l] = lv.db2[l] = nullptr; lv.H[l] = lv.W[l] = 1;
    }
  }
  for (int l = L; l <= kMaxLevels; ++l) lv.cta0[l] = cta;
  lv.L = L; lv.B = B; lv.C = C; lv.need_dx = dx != nullptr;
  if (cta == 0) return cudaSuccess;
  for (int l = 0; l < L; ++l)
    if ((long long)B * H[l] * W[l] >= (1ll << 31)) return cudaErrorNotSupported;
  cudaError_t e;
  if (dtype == 0) {
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(gate_conv_bwd<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    gate_conv_bwd<float><<<cta, kThreads, smem, stream>>>(lv);
  } else {
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(gate_conv_bwd<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    gate_conv_bwd<__nv_bfloat16><<<cta, kThreads, smem, stream>>>(lv);
  }
  return cudaGetLastError();
}

}  // namespace arfe
