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
