// Index rules and small helpers shared by the AR-FPN kernels (fpn.cu: NCHW and
// generic paths; fpn_nhwc.cu: channels-last vector paths).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "launch.h"

namespace arfe {
namespace fpn {

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

// Map sizes are validated < 32768 at the C ABI, so the products fit 32 bits.
__device__ __forceinline__ int pool_start(int i, int in, int out) {
  return (int)(((unsigned)i * (unsigned)in) / (unsigned)out);
}
__device__ __forceinline__ int pool_end(int i, int in, int out) {
  return (int)(((unsigned)(i + 1) * (unsigned)in + (unsigned)out - 1u) / (unsigned)out);
}
__device__ __forceinline__ int nearest_src(int d, int in, int out) {
  const float scale = __fdiv_rn((float)in, (float)out);
  const int s = (int)floorf(__fmul_rn((float)d, scale));
  return s < in - 1 ? s : in - 1;
}


__device__ __forceinline__ float gate_value(float a, float b) {
  return tanhf(fmaxf(a, 0.f)) + tanhf(fmaxf(b, 0.f));
}

// Level rows y (of a map H tall) whose nearest source in a map Hr tall is Y,
// i.e. nearest_src(y, Hr, H) == Y: a contiguous range [ya, yb).
__device__ __forceinline__ void dst_range(int Y, int Hr, int H, int& ya, int& yb) {
  const float s = (float)H / (float)Hr;
  ya = (int)floorf((float)Y * s) - 1; if (ya < 0) ya = 0;
  yb = (int)ceilf((float)(Y + 1) * s) + 1; if (yb > H) yb = H;
  while (ya < yb && nearest_src(ya, Hr, H) != Y) ++ya;
  while (yb > ya && nearest_src(yb - 1, Hr, H) != Y) --yb;
}

}  // namespace fpn
}  // namespace arfe
