// AR-RFF fusion gate: out = ori + ori * (a + b)
//   MultiBBoxHead.forward, mmdet/models/roi_heads/bbox_heads/multirois_bbox_head.py:175,182
// `ori` is read in place from the concatenated [K, 3C, PH, PW] extraction
// output (row stride = ori_stride elements), so the reference's channel slice
// copy never materialises.  Pure streaming: 3 reads + 1 write (forward),
// 4 reads + 2 writes (backward; da == db so one buffer is written).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.h"

namespace arfe {
namespace {

constexpr int kThreads = 256;

template <typename T, int V> struct Vec;
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<float, 1> { using type = float; };
template <> struct Vec<__nv_bfloat16, 8> { using type = uint4; };
template <> struct Vec<__nv_bfloat16, 1> { using type = __nv_bfloat16; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[V]) {
  using VT = typename Vec<T, V>::type;
  VT raw = *reinterpret_cast<const VT*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if constexpr (sizeof(T) == 4) f[i] = (float)e[i];
    else f[i] = __bfloat162float(e[i]);
  }
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[V]) {
  using VT = typename Vec<T, V>::type;
  VT raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if constexpr (sizeof(T) == 4) e[i] = (T)f[i];
    else e[i] = __float2bfloat16_rn(f[i]);
  }
  *reinterpret_cast<VT*>(p) = raw;
}

// nv = n / V vectors per RoI, total = K * nv
template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
gate_fwd(const T* __restrict__ ori, int64_t ori_stride, const T* __restrict__ a,
         const T* __restrict__ b, T* __restrict__ out, int64_t total, uint32_t nv,
         int64_t n) {
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += step) {
    const int64_t k = i / nv;
    const int64_t j = (i - k * nv) * V;
    float fo[V], fa[V], fb[V], r[V];
    load_vec<T, V>(ori + k * ori_stride + j, fo);
    load_vec<T, V>(a + k * n + j, fa);
    load_vec<T, V>(b + k * n + j, fb);
#pragma unroll
    for (int q = 0; q < V; ++q) r[q] = fmaf(fo[q], fa[q] + fb[q], fo[q]);
    store_vec<T, V>(out + k * n + j, r);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
gate_bwd(const T* __restrict__ g, const T* __restrict__ ori, int64_t ori_stride,
         const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ d_ori,
         int64_t d_ori_stride, T* __restrict__ d_ab, int64_t total, uint32_t nv, int64_t n) {
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += step) {
    const int64_t k = i / nv;
    const int64_t j = (i - k * nv) * V;
    float fg[V], fo[V], fa[V], fb[V], r0[V], r1[V];
    load_vec<T, V>(g + k * n + j, fg);
    load_vec<T, V>(ori + k * ori_stride + j, fo);
    load_vec<T, V>(a + k * n + j, fa);
    load_vec<T, V>(b + k * n + j, fb);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      r0[q] = fg[q] * (1.0f + (fa[q] + fb[q]));
      r1[q] = fg[q] * fo[q];
    }
    store_vec<T, V>(d_ori + k * d_ori_stride + j, r0);
    store_vec<T, V>(d_ab + k * n + j, r1);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int grid_for(int64_t total) {
  int64_t blocks = (total + kThreads - 1) / kThreads;
  const int64_t cap = 148 * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

cudaError_t launch_rff_gate_forward(const void* ori, int64_t ori_stride, const void* a,
                                    const void* b, void* out, int64_t K, int64_t n,
                                    int dtype, cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const bool vec = (n % V == 0) && (ori_stride % V == 0) && aligned16(ori) &&
                   aligned16(a) && aligned16(b) && aligned16(out);
  if (dtype == 0) {
    auto o = (const float*)ori; auto pa = (const float*)a; auto pb = (const float*)b; auto po = (float*)out;
    if (vec) gate_fwd<float, 4><<<grid_for(K * n / 4), kThreads, 0, stream>>>(o, ori_stride, pa, pb, po, K * (n / 4), (uint32_t)(n / 4), n);
    else gate_fwd<float, 1><<<grid_for(K * n), kThreads, 0, stream>>>(o, ori_stride, pa, pb, po, K * n, (uint32_t)n, n);
  } else {
    auto o = (const __nv_bfloat16*)ori; auto pa = (const __nv_bfloat16*)a; auto pb = (const __nv_bfloat16*)b; auto po = (__nv_bfloat16*)out;
    if (vec) gate_fwd<__nv_bfloat16, 8><<<grid_for(K * n / 8), kThreads, 0, stream>>>(o, ori_stride, pa, pb, po, K * (n / 8), (uint32_t)(n / 8), n);
    else gate_fwd<__nv_bfloat16, 1><<<grid_for(K * n), kThreads, 0, stream>>>(o, ori_stride, pa, pb, po, K * n, (uint32_t)n, n);
  }
  return cudaGetLastError();
}

cudaError_t launch_rff_gate_backward(const void* g, const void* ori, int64_t ori_stride,
                                     const void* a, const void* b, void* d_ori,
                                     int64_t d_ori_stride, void* d_ab, int64_t K, int64_t n,
                                     int dtype, cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const bool vec = (n % V == 0) && (ori_stride % V == 0) && (d_ori_stride % V == 0) &&
                   aligned16(g) && aligned16(ori) && aligned16(a) && aligned16(b) &&
                   aligned16(d_ori) && aligned16(d_ab);
  if (dtype == 0) {
    auto pg = (const float*)g; auto o = (const float*)ori; auto pa = (const float*)a; auto pb = (const float*)b;
    auto d0 = (float*)d_ori; auto d1 = (float*)d_ab;
    if (vec) gate_bwd<float, 4><<<grid_for(K * n / 4), kThreads, 0, stream>>>(pg, o, ori_stride, pa, pb, d0, d_ori_stride, d1, K * (n / 4), (uint32_t)(n / 4), n);
    else gate_bwd<float, 1><<<grid_for(K * n), kThreads, 0, stream>>>(pg, o, ori_stride, pa, pb, d0, d_ori_stride, d1, K * n, (uint32_t)n, n);
  } else {
    auto pg = (const __nv_bfloat16*)g; auto o = (const __nv_bfloat16*)ori; auto pa = (const __nv_bfloat16*)a; auto pb = (const __nv_bfloat16*)b;
    auto d0 = (__nv_bfloat16*)d_ori; auto d1 = (__nv_bfloat16*)d_ab;
    if (vec) gate_bwd<__nv_bfloat16, 8><<<grid_for(K * n / 8), kThreads, 0, stream>>>(pg, o, ori_stride, pa, pb, d0, d_ori_stride, d1, K * (n / 8), (uint32_t)(n / 8), n);
    else gate_bwd<__nv_bfloat16, 1><<<grid_for(K * n), kThreads, 0, stream>>>(pg, o, ori_stride, pa, pb, d0, d_ori_stride, d1, K * n, (uint32_t)n, n);
  }
  return cudaGetLastError();
}

}  // namespace arfe

// ---------------------------------------------------------------------------
// AR-RFF, softmax-over-regions variant (the fusion of the paper's figure; in the
// reference it is the commented block multirois_bbox_head.py:187-197):
//   ws  = softmax(logits, dim = 1)            logits [K, 3, PH, PW]
//   out = r0 * ws[:, 0] + r1 * ws[:, 1] + r2 * ws[:, 2]      r_j [K, C, PH, PW]
// backward: d r_j = d out * ws_j;  d logit_j = ws_j * (s_j - sum_i ws_i s_i) with
// s_j = sum_c d out * r_j  (a reduction over the C channels of one bin).
// Element (k, bin, c) of a region / out tensor sits at k * ks + bin * bs + c * cs:
// channels-last tensors (cs == 1; also channel slices of the concatenated tensor) take
// the warp-per-bin kernels (lanes over channels, 128-bit accesses, warp-shuffle
// reduction), NCHW tensors (bs == 1) the thread-per-bin kernels (lanes over bins).
namespace arfe {
namespace {

struct FuseAddr { int64_t ks, bs, cs; };

__device__ __forceinline__ void softmax3(float l0, float l1, float l2, float (&w)[3]) {
  const float m = fmaxf(l0, fmaxf(l1, l2));
  const float e0 = expf(l0 - m), e1 = expf(l1 - m), e2 = expf(l2 - m);
  const float s = e0 + e1 + e2;
  w[0] = e0 / s; w[1] = e1 / s; w[2] = e2 / s;
}
template <typename T> __device__ __forceinline__ float ldsc(const T* p) {
  if constexpr (sizeof(T) == 4) return *p; else return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stsc(T* p, float v) {
  if constexpr (sizeof(T) == 4) *p = v; else *p = __float2bfloat16_rn(v);
}

// warp == one (RoI, bin) row; V channels per lane step (channels-last: cs == 1)
template <typename T, int V, bool kBackward>
__global__ void __launch_bounds__(kThreads)
softmax_fuse_cl(const T* __restrict__ r0, const T* __restrict__ r1, const T* __restrict__ r2, FuseAddr ra,
                const T* __restrict__ logits, int64_t lks, int64_t lrs, int64_t lbs,
                T* __restrict__ out, const T* __restrict__ dout, FuseAddr oa,
                T* __restrict__ d0, T* __restrict__ d1, T* __restrict__ d2, T* __restrict__ dlogits,
                int64_t rows, int PP, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t k = row / PP;
  const int bin = (int)(row - k * PP);
  const T* lg = logits + k * lks + (int64_t)bin * lbs;
  float w[3];
  softmax3(ldsc(lg), ldsc(lg + lrs), ldsc(lg + 2 * lrs), w);
  const int64_t ro = k * ra.ks + (int64_t)bin * ra.bs, oo = k * oa.ks + (int64_t)bin * oa.bs;
  float s[3] = {0.f, 0.f, 0.f};
  for (int c = lane * V; c < C; c += 32 * V) {
    float a[V], b[V], d[V];
    load_vec<T, V>(r0 + ro + c, a);
    load_vec<T, V>(r1 + ro + c, b);
    load_vec<T, V>(r2 + ro + c, d);
    if constexpr (!kBackward) {
      float o[V];
#pragma unroll
      for (int q = 0; q < V; ++q) o[q] = fmaf(d[q], w[2], fmaf(b[q], w[1], a[q] * w[0]));
      store_vec<T, V>(out + oo + c, o);
    } else {
      float g[V], x0[V], x1[V], x2[V];
      load_vec<T, V>(dout + oo + c, g);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        s[0] = fmaf(g[q], a[q], s[0]); s[1] = fmaf(g[q], b[q], s[1]); s[2] = fmaf(g[q], d[q], s[2]);
        x0[q] = g[q] * w[0]; x1[q] = g[q] * w[1]; x2[q] = g[q] * w[2];
      }
      store_vec<T, V>(d0 + oo + c, x0);
      store_vec<T, V>(d1 + oo + c, x1);
      store_vec<T, V>(d2 + oo + c, x2);
    }
  }
  if constexpr (kBackward) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int dd = 16; dd > 0; dd >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], dd);
    if (lane < 3) {
      const float mean = w[0] * s[0] + w[1] * s[1] + w[2] * s[2];
      const float sj = lane == 0 ? s[0] : (lane == 1 ? s[1] : s[2]);
      const float wj = lane == 0 ? w[0] : (lane == 1 ? w[1] : w[2]);
      stsc(dlogits + k * lks + (int64_t)bin * lbs + lane * lrs, wj * (sj - mean));
    }
  }
}

// thread == one (RoI, bin); loops over the channels (NCHW: consecutive threads, consecutive bins)
template <typename T, bool kBackward>
__global__ void __launch_bounds__(kThreads)
softmax_fuse_any(const T* __restrict__ r0, const T* __restrict__ r1, const T* __restrict__ r2, FuseAddr ra,
                 const T* __restrict__ logits, int64_t lks, int64_t lrs, int64_t lbs,
                 T* __restrict__ out, const T* __restrict__ dout, FuseAddr oa,
                 T* __restrict__ d0, T* __restrict__ d1, T* __restrict__ d2, T* __restrict__ dlogits,
                 int64_t rows, int PP, int C) {
  const int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (row >= rows) return;
  const int64_t k = row / PP;
  const int bin = (int)(row - k * PP);
  const T* lg = logits + k * lks + (int64_t)bin * lbs;
  float w[3];
  softmax3(ldsc(lg), ldsc(lg + lrs), ldsc(lg + 2 * lrs), w);
  const int64_t ro = k * ra.ks + (int64_t)bin * ra.bs, oo = k * oa.ks + (int64_t)bin * oa.bs;
  float s[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) {
    const float a = ldsc(r0 + ro + c * ra.cs), b = ldsc(r1 + ro + c * ra.cs), d = ldsc(r2 + ro + c * ra.cs);
    if constexpr (!kBackward) {
      stsc(out + oo + c * oa.cs, fmaf(d, w[2], fmaf(b, w[1], a * w[0])));
    } else {
      const float g = ldsc(dout + oo + c * oa.cs);
      s[0] = fmaf(g, a, s[0]); s[1] = fmaf(g, b, s[1]); s[2] = fmaf(g, d, s[2]);
      stsc(d0 + oo + c * oa.cs, g * w[0]);
      stsc(d1 + oo + c * oa.cs, g * w[1]);
      stsc(d2 + oo + c * oa.cs, g * w[2]);
    }
  }
  if constexpr (kBackward) {
    const float mean = w[0] * s[0] + w[1] * s[1] + w[2] * s[2];
    T* dl = dlogits + k * lks + (int64_t)bin * lbs;
    stsc(dl, w[0] * (s[0] - mean));
    stsc(dl + lrs, w[1] * (s[1] - mean));
    stsc(dl + 2 * lrs, w[2] * (s[2] - mean));
  }
}

template <typename T, bool kBackward>
cudaError_t launch_softmax_fuse_t(const void* const* reg, const int64_t* rstr, const void* logits,
                                  const int64_t* lstr, void* out, const void* dout, const int64_t* ostr,
                                  void* const* dreg, void* dlogits, int64_t K, int PP, int C,
                                  cudaStream_t stream) {
  const int64_t rows = K * PP;
  if (rows == 0) return cudaSuccess;
  constexpr int V = sizeof(T) == 4 ? 4 : 8;
  const FuseAddr ra{rstr[0], rstr[1], rstr[2]}, oa{ostr[0], ostr[1], ostr[2]};
  auto r0 = (const T*)reg[0]; auto r1 = (const T*)reg[1]; auto r2 = (const T*)reg[2];
  T* d0 = kBackward ? (T*)dreg[0] : nullptr; T* d1 = kBackward ? (T*)dreg[1] : nullptr; T* d2 = kBackward ? (T*)dreg[2] : nullptr;
  const bool cl = ra.cs == 1 && oa.cs == 1;
  if (cl) {
    bool vec = C % V == 0 && ra.ks % V == 0 && ra.bs % V == 0 && oa.ks % V == 0 && oa.bs % V == 0 &&
               aligned16(r0) && aligned16(r1) && aligned16(r2) && aligned16(kBackward ? dout : out);
    if (kBackward) vec = vec && aligned16(d0) && aligned16(d1) && aligned16(d2);
    const unsigned grid = (unsigned)((rows + kThreads / 32 - 1) / (kThreads / 32));
    if (vec) softmax_fuse_cl<T, V, kBackward><<<grid, kThreads, 0, stream>>>(r0, r1, r2, ra, (const T*)logits, lstr[0], lstr[1], lstr[2], (T*)out, (const T*)dout, oa, d0, d1, d2, (T*)dlogits, rows, PP, C);
    else softmax_fuse_cl<T, 1, kBackward><<<grid, kThreads, 0, stream>>>(r0, r1, r2, ra, (const T*)logits, lstr[0], lstr[1], lstr[2], (T*)out, (const T*)dout, oa, d0, d1, d2, (T*)dlogits, rows, PP, C);
  } else {
    const unsigned grid = (unsigned)((rows + kThreads - 1) / kThreads);
    softmax_fuse_any<T, kBackward><<<grid, kThreads, 0, stream>>>(r0, r1, r2, ra, (const T*)logits, lstr[0], lstr[1], lstr[2], (T*)out, (const T*)dout, oa, d0, d1, d2, (T*)dlogits, rows, PP, C);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_rff_softmax_fuse(int backward, const void* const* reg, const int64_t* rstr, const void* logits,
                                    const int64_t* lstr, void* out, const void* dout, const int64_t* ostr,
                                    void* const* dreg, void* dlogits, int64_t K, int PP, int C, int dtype,
                                    cudaStream_t stream) {
  if (dtype == 0)
    return backward ? launch_softmax_fuse_t<float, true>(reg, rstr, logits, lstr, out, dout, ostr, dreg, dlogits, K, PP, C, stream)
                    : launch_softmax_fuse_t<float, false>(reg, rstr, logits, lstr, out, dout, ostr, dreg, dlogits, K, PP, C, stream);
  return backward ? launch_softmax_fuse_t<__nv_bfloat16, true>(reg, rstr, logits, lstr, out, dout, ostr, dreg, dlogits, K, PP, C, stream)
                  : launch_softmax_fuse_t<__nv_bfloat16, false>(reg, rstr, logits, lstr, out, dout, ostr, dreg, dlogits, K, PP, C, stream);
}

}  // namespace arfe
