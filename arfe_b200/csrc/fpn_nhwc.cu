// AR-FPN kernels, channels-last (NHWC) vector paths: lanes == channels, 128-bit
// accesses, every access coalesced over the channel axis.
//
//   gather fwd : thread == (refine pixel, 4|8 channels): window max of the
//                pooled levels + nearest taps of the others, summed in level
//                order, / L.   (wfpn_dual_spatial.py:102-113)
//   gather bwd : same thread shape routes g/L to the argmax cell of every
//                exact-ratio pooling window and writes the refine level.
//   apply fwd  : ONE WARP per refine pixel walks the level pixels that read it
//                (4x4, 2x2, 1, ...): the gate (two tanh) is evaluated once per
//                pixel, bsf is held in registers.   (wfpn_dual_spatial.py:118-135)
//   apply bwd  : same walk, fused: d(gate sum) = sum_c dout*bsf by warp shuffle
//                -> dg1, dg2; dbsf += dout*gate in registers.  dout is read
//                exactly once, nothing is atomic, no scratch buffer.
// Index rules as in fpn.cu.  Requires C % (4|8) == 0 and 16-byte aligned
// tensors; anything else takes the generic kernels in fpn.cu.
#include <stdlib.h>
#include <type_traits>

#include "fpn_common.cuh"
#include "tma.cuh"

namespace arfe {
using namespace fpn;
namespace {

template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int n = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int n = 8; };

template <typename T>
__device__ __forceinline__ void ldv(const T* __restrict__ p, float (&f)[Vec<T>::n]) {
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
// V consecutive channels from shared memory
template <typename T>
__device__ __forceinline__ void lds_vec(const unsigned char* p, float (&f)[Vec<T>::n]) {
  if constexpr (sizeof(T) == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
}
template <typename T>
__device__ __forceinline__ void stv(T* __restrict__ p, const float (&f)[Vec<T>::n]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  } else {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
}

// f[] += V consecutive fp32 values at a (a != NULL)
template <int V>
__device__ __forceinline__ void add_f32(const float* __restrict__ a, float (&f)[V]) {
#pragma unroll
  for (int u = 0; u < V; u += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(a + u));
    f[u] += t.x; f[u + 1] += t.y; f[u + 2] += t.z; f[u + 3] += t.w;
  }
}

struct Exact { int s[kMaxLevels]; };  // pooling ratio of level l (< refine) if exact, else 0

// ---------------------------------------------------------------- gather fwd
// argmax layout here: [l][B][Hr][Wr][C] (follows the channels-last layout).
template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_fwd_cl(const FpnParams p, const Exact ex) {
  constexpr int V = Vec<T>::n;
  const int Hr = p.Hr, Wr = p.Wr, C = p.C, CV = C / V;
  const size_t total = (size_t)p.B * Hr * Wr * CV;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % CV);
  size_t r = i / CV;
  const int X = (int)(r % Wr); r /= Wr;
  const int Y = (int)(r % Hr);
  const int b = (int)(r / Hr);
  const int c = cv * V;
  float acc[V];
#pragma unroll
  for (int u = 0; u < V; ++u) acc[u] = 0.f;
  for (int l = 0; l < p.L; ++l) {
    const T* __restrict__ f = static_cast<const T*>(p.feats[l]);
    const int H = p.H[l], W = p.W[l];
    float v[V];
    if (l < p.refine_level) {
      int arg[V];
#pragma unroll
      for (int u = 0; u < V; ++u) { v[u] = -CUDART_INF_F; arg[u] = 0; }
      const int s = ex.s[l];
      if (s == 4 || s == 2) {
        // exact ratio: static window, a whole window row of loads in flight
        const T* __restrict__ w0 = f + (((size_t)b * H + (size_t)s * Y) * W + (size_t)s * X) * C + c;
        if (s == 4) {
#pragma unroll
          for (int dy = 0; dy < 4; ++dy) {
            float t[4][V];
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) ldv<T>(w0 + ((size_t)dy * W + dx) * C, t[dx]);
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
#pragma unroll
              for (int u = 0; u < V; ++u)
                if (t[dx][u] > v[u] || t[dx][u] != t[dx][u]) { v[u] = t[dx][u]; arg[u] = dy * 4 + dx; }
          }
        } else {
          float t[4][V];
#pragma unroll
          for (int k = 0; k < 4; ++k) ldv<T>(w0 + ((size_t)(k >> 1) * W + (k & 1)) * C, t[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int u = 0; u < V; ++u)
              if (t[k][u] > v[u] || t[k][u] != t[k][u]) { v[u] = t[k][u]; arg[u] = k; }
        }
      } else {
        const int y0 = pool_start(Y, H, Hr), y1 = pool_end(Y, H, Hr);
        const int x0 = pool_start(X, W, Wr), x1 = pool_end(X, W, Wr);
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) {
            float t[V];
            ldv<T>(f + (((size_t)b * H + y) * W + x) * C + c, t);
            const int pos = (y - y0) * (x1 - x0) + (x - x0);
#pragma unroll
            for (int u = 0; u < V; ++u)
              if (t[u] > v[u] || t[u] != t[u]) { v[u] = t[u]; arg[u] = pos; }
          }
      }
      if (p.argmax) {
        uint8_t* a = p.argmax + ((((size_t)l * p.B + b) * Hr + Y) * Wr + X) * C + c;
#pragma unroll
        for (int u = 0; u < V; u += 4)
          *reinterpret_cast<uchar4*>(a + u) =
              make_uchar4((uint8_t)arg[u], (uint8_t)arg[u + 1], (uint8_t)arg[u + 2], (uint8_t)arg[u + 3]);
      }
    } else {
      ldv<T>(f + (((size_t)b * H + nearest_src(Y, H, Hr)) * W + nearest_src(X, W, Wr)) * C + c, v);
    }
#pragma unroll
    for (int u = 0; u < V; ++u) acc[u] = __fadd_rn(acc[u], v[u]);
  }
#pragma unroll
  for (int u = 0; u < V; ++u) acc[u] = __fdiv_rn(acc[u], (float)p.L);
  stv<T>(static_cast<T*>(p.gathered) + i * V, acc);
}

// ---------------------------------------------------------------- gather bwd
// Exact-ratio pooled levels + the refine level, one thread per (refine px, vec).
template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_bwd_cl(const FpnParams p, const Exact ex) {
  constexpr int V = Vec<T>::n;
  const int Hr = p.Hr, Wr = p.Wr, C = p.C, CV = C / V;
  const size_t total = (size_t)p.B * Hr * Wr * CV;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % CV);
  size_t r = i / CV;
  const int X = (int)(r % Wr); r /= Wr;
  const int Y = (int)(r % Hr);
  const int b = (int)(r / Hr);
  const int c = cv * V;
  float g[V];
  ldv<T>(static_cast<const T*>(p.gathered) + i * V, g);
#pragma unroll
  for (int u = 0; u < V; ++u) g[u] = __fdiv_rn(g[u], (float)p.L);
  for (int l = 0; l < p.refine_level; ++l) {
    const int s = ex.s[l];
    if (s == 0) continue;
    const int H = p.H[l], W = p.W[l];
    const uint8_t* a = p.argmax + ((((size_t)l * p.B + b) * Hr + Y) * Wr + X) * C + c;
    int arg[V];
#pragma unroll
    for (int u = 0; u < V; u += 4) {
      const uchar4 q = *reinterpret_cast<const uchar4*>(a + u);
      arg[u] = q.x; arg[u + 1] = q.y; arg[u + 2] = q.z; arg[u + 3] = q.w;
    }
    T* __restrict__ o = static_cast<T*>(p.outs[l]);
    const float* __restrict__ add = p.addend[l];
    for (int dy = 0; dy < s; ++dy)
      for (int dx = 0; dx < s; ++dx) {
        const int pos = dy * s + dx;
        const size_t at = (((size_t)b * H + (size_t)s * Y + dy) * W + (size_t)s * X + dx) * C + c;
        float v[V];
#pragma unroll
        for (int u = 0; u < V; ++u) v[u] = (arg[u] == pos) ? g[u] : 0.f;
        if (add) add_f32<V>(add + at, v);
        stv<T>(o + at, v);
      }
  }
  if (p.addend[p.refine_level]) add_f32<V>(p.addend[p.refine_level] + i * V, g);
  stv<T>(static_cast<T*>(p.outs[p.refine_level]) + i * V, g);
}

// Levels above the refine level (smaller maps, nearest-upsampled in the forward):
// thread == (level pixel, vector): sum of d(gathered) over the refine pixels whose
// nearest source it is.
struct UpLevels { size_t start[kMaxLevels + 1]; int level[kMaxLevels]; int n; };

template <typename T>
__global__ void __launch_bounds__(kThreads)
gather_bwd_up_cl(const FpnParams p, const UpLevels ul) {
  constexpr int V = Vec<T>::n;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= ul.start[ul.n]) return;
  int j = 0;
  while (j + 1 < ul.n && i >= ul.start[j + 1]) ++j;
  const int l = ul.level[j];
  const int H = p.H[l], W = p.W[l], C = p.C, CV = C / V, Hr = p.Hr, Wr = p.Wr;
  size_t r = i - ul.start[j];
  const int cv = (int)(r % CV); r /= CV;
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H);
  const int b = (int)(r / H);
  const int c = cv * V;
  // refine rows Y with nearest_src(Y: H <- Hr) == y
  const float sy = (float)Hr / (float)H, sx = (float)Wr / (float)W;
  int Ya = (int)floorf((float)y * sy) - 1; if (Ya < 0) Ya = 0;
  int Yb = (int)ceilf((float)(y + 1) * sy) + 1; if (Yb > Hr) Yb = Hr;
  int Xa = (int)floorf((float)x * sx) - 1; if (Xa < 0) Xa = 0;
  int Xb = (int)ceilf((float)(x + 1) * sx) + 1; if (Xb > Wr) Xb = Wr;
  while (Ya < Yb && nearest_src(Ya, H, Hr) != y) ++Ya;
  while (Yb > Ya && nearest_src(Yb - 1, H, Hr) != y) --Yb;
  while (Xa < Xb && nearest_src(Xa, W, Wr) != x) ++Xa;
  while (Xb > Xa && nearest_src(Xb - 1, W, Wr) != x) --Xb;
  float g[V];
#pragma unroll
  for (int u = 0; u < V; ++u) g[u] = 0.f;
  const T* __restrict__ d = static_cast<const T*>(p.gathered);
  if (Yb - Ya <= 4 && Xb - Xa <= 4) {
    // the usual case (ratios up to 4): a whole window row of loads in flight instead of
    // one DRAM latency per refine pixel; same summation order as the loop below
    const T* __restrict__ w0 = d + (((size_t)b * Hr + Ya) * Wr + Xa) * C + c;
    const int ny = Yb - Ya, nx = Xb - Xa;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      if (dy >= ny) break;
      float t[4][V];
#pragma unroll
      for (int dx = 0; dx < 4; ++dx)
        if (dx < nx) ldv<T>(w0 + ((size_t)dy * Wr + dx) * C, t[dx]);
#pragma unroll
      for (int dx = 0; dx < 4; ++dx)
        if (dx < nx) {
#pragma unroll
          for (int u = 0; u < V; ++u) g[u] += t[dx][u];
        }
    }
  } else {
    for (int Y = Ya; Y < Yb; ++Y)
      for (int X = Xa; X < Xb; ++X) {
        float t[V];
        ldv<T>(d + (((size_t)b * Hr + Y) * Wr + X) * C + c, t);
#pragma unroll
        for (int u = 0; u < V; ++u) g[u] += t[u];
      }
  }
#pragma unroll
  for (int u = 0; u < V; ++u) g[u] = __fdiv_rn(g[u], (float)p.L);
  if (p.addend[l]) add_f32<V>(p.addend[l] + (i - ul.start[j]) * V, g);
  stv<T>(static_cast<T*>(p.outs[l]) + (i - ul.start[j]) * V, g);
}

// ------------------------------------------------------------ apply fwd/bwd
// One warp per refine pixel (b, Y, X); NV vectors per lane cover C channels.
template <typename T, int NV, bool kBackward, int kOcc>
__global__ void __launch_bounds__(kThreads, kOcc)
apply_cl(const FpnParams p) {
  constexpr int V = Vec<T>::n;
  const int Hr = p.Hr, Wr = p.Wr, C = p.C;
  const int lane = threadIdx.x & 31;
  const size_t wid = (size_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (wid >= (size_t)p.B * Hr * Wr) return;
  const int X = (int)(wid % Wr);
  const int Y = (int)((wid / Wr) % Hr);
  const int b = (int)(wid / ((size_t)Wr * Hr));
  // this lane's channels: c = (v * 32 + lane) * V, v < NV
  float bs[NV][V], db[NV][V];
  bool on[NV];
  const T* __restrict__ bsf = static_cast<const T*>(p.bsf) + (((size_t)b * Hr + Y) * Wr + X) * C;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * V;
    on[v] = c < C;
    if (on[v]) ldv<T>(bsf + c, bs[v]);
    else {
#pragma unroll
      for (int u = 0; u < V; ++u) bs[v][u] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < V; ++u) db[v][u] = 0.f;
  }
  for (int l = 0; l < p.L; ++l) {
    const int H = p.H[l], W = p.W[l];
    int ya, yb, xa, xb;
    dst_range(Y, Hr, H, ya, yb);
    dst_range(X, Wr, W, xa, xb);
    const T* __restrict__ xin = static_cast<const T*>(p.feats[l]);  // x_l (fwd) / dout_l (bwd)
    const T* __restrict__ q1 = static_cast<const T*>(p.g1[l]);
    const T* __restrict__ q2 = static_cast<const T*>(p.g2[l]);
    // NPX pixels of a level row at a time: all their loads (gate maps + NV
    // vectors each) are issued before the first use, so a warp pays the DRAM
    // latency once per group instead of once per pixel (a warp visits ~22 pixels)
    auto group = [&](auto npx_tag, int y, int x) {
      constexpr int NPX = decltype(npx_tag)::value;
      const size_t pix0 = ((size_t)b * H + y) * W + x;
      float a1[NPX], a2[NPX], f[NPX][NV][V];
#pragma unroll
      for (int n = 0; n < NPX; ++n) {
        a1[n] = ldf(q1 + pix0 + n);
        a2[n] = ldf(q2 + pix0 + n);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (on[v]) ldv<T>(xin + (pix0 + n) * C + (v * 32 + lane) * V, f[n][v]);
          else {
#pragma unroll
            for (int u = 0; u < V; ++u) f[n][v][u] = 0.f;
          }
        }
      }
#pragma unroll
      for (int n = 0; n < NPX; ++n) {
        const size_t pix = pix0 + n;
        const float t1 = tanhf(fmaxf(a1[n], 0.f)), t2 = tanhf(fmaxf(a2[n], 0.f));
        const float gate = t1 + t2;
        if (!kBackward) {
          T* __restrict__ out = static_cast<T*>(p.outs[l]);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (!on[v]) continue;
#pragma unroll
            for (int u = 0; u < V; ++u) f[n][v][u] = fmaf(bs[v][u], gate, f[n][v][u]);
            stv<T>(out + pix * C + (v * 32 + lane) * V, f[n][v]);
          }
        } else {
          float s = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
#pragma unroll
            for (int u = 0; u < V; ++u) {
              s = fmaf(f[n][v][u], bs[v][u], s);
              db[v][u] = fmaf(f[n][v][u], gate, db[v][u]);
            }
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
          if (lane == 0) {
            p.dg1[l][pix] = a1[n] > 0.f ? s * (1.f - t1 * t1) : 0.f;
            p.dg2[l][pix] = a2[n] > 0.f ? s * (1.f - t2 * t2) : 0.f;
          }
        }
      }
    };
    if constexpr (!kBackward) {
      // forward 67 -> 62 us with grouped loads
      for (int y = ya; y < yb; ++y) {
        int x = xa;
        for (; x + 4 <= xb; x += 4) group(std::integral_constant<int, 4>{}, y, x);
        for (; x + 2 <= xb; x += 2) group(std::integral_constant<int, 2>{}, y, x);
        for (; x < xb; ++x) group(std::integral_constant<int, 1>{}, y, x);
      }
    } else {
      // backward: pixel by pixel (groups of 4 / 2 / 1 through the lambda measured
      // 77 / 70 / 83 us against 62 us for this loop: registers, occupancy)
      for (int y = ya; y < yb; ++y)
        for (int x = xa; x < xb; ++x) {
          const size_t pix = ((size_t)b * H + y) * W + x;
          const float a1 = ldf(q1 + pix), a2 = ldf(q2 + pix);
          const float t1 = tanhf(fmaxf(a1, 0.f)), t2 = tanhf(fmaxf(a2, 0.f));
          const float gate = t1 + t2;
          float s = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (!on[v]) continue;
            const int c = (v * 32 + lane) * V;
            float f[V];
            ldv<T>(xin + pix * C + c, f);
#pragma unroll
            for (int u = 0; u < V; ++u) {
              s = fmaf(f[u], bs[v][u], s);
              db[v][u] = fmaf(f[u], gate, db[v][u]);
            }
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
          if (lane == 0) {
            p.dg1[l][pix] = a1 > 0.f ? s * (1.f - t1 * t1) : 0.f;
            p.dg2[l][pix] = a2 > 0.f ? s * (1.f - t2 * t2) : 0.f;
          }
        }
    }
  }
  if (kBackward) {
    float* __restrict__ o = p.dbsf + (((size_t)b * Hr + Y) * Wr + X) * C;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!on[v]) continue;
      const int c = (v * 32 + lane) * V;
#pragma unroll
      for (int u = 0; u < V; u += 4)
        *reinterpret_cast<float4*>(o + c + u) = make_float4(db[v][u], db[v][u + 1], db[v][u + 2], db[v][u + 3]);
    }
  }
}

// ------------------------------------------------- fused AR-FPN backward
// apply backward + gather backward in ONE pass over the incoming gradient pyramid.
// x_l feeds the gather (wfpn_dual_spatial.py:102-113) and the gated residual (:135), so
//   d x_l = d out_l + (gather's routed gradient),
// and d out_l is also what the gate / bsf gradients are made of.  The warp of refine pixel
// (b, Y, X) walks the level pixels that read it -- for the pooled levels (exact ratio s)
// that is exactly its s x s pooling window, for the refine level the pixel itself, for the
// upsampled levels the level pixel whose nearest destination it is -- so every d out
// element is read once and every d x element written once: the separate kernels read the
// pyramid twice.  Arithmetic and summation order are those of apply_cl<kBackward> and
// gather_bwd_cl / gather_bwd_up_cl (the results are bit-identical to the two-call path).
// TD = element type of d out (fp32 straight from the RoI backward, or T).
template <typename TD, int V>
__device__ __forceinline__ void ldv_as_f32(const TD* __restrict__ p, float (&f)[V]) {
  if constexpr (sizeof(TD) == 4) {
#pragma unroll
    for (int u = 0; u < V; u += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + u));
      f[u] = v.x; f[u + 1] = v.y; f[u + 2] = v.z; f[u + 3] = v.w;
    }
  } else {
    static_assert(V == 8, "bf16 vectors are 8 wide");
    ldv<TD>(p, f);
  }
}

template <typename T, typename TD, int NV>
__global__ void __launch_bounds__(kThreads, (Vec<T>::n * NV > 8 ? 3 : 4))
fpn_bwd_fused_cl(const FpnParams p, const Exact ex) {
  constexpr int V = Vec<T>::n;
  const int Hr = p.Hr, Wr = p.Wr, C = p.C;
  const int lane = threadIdx.x & 31;
  const size_t wid = (size_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (wid >= (size_t)p.B * Hr * Wr) return;
  const int X = (int)(wid % Wr);
  const int Y = (int)((wid / Wr) % Hr);
  const int b = (int)(wid / ((size_t)Wr * Hr));
  const float Lf = (float)p.L;
  float bs[NV][V], db[NV][V], gg[NV][V];
  bool on[NV];
  const size_t rpix = ((size_t)b * Hr + Y) * Wr + X;
  const T* __restrict__ dga = static_cast<const T*>(p.gathered);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * V;
    on[v] = c < C;
#pragma unroll
    for (int u = 0; u < V; ++u) { bs[v][u] = 0.f; db[v][u] = 0.f; gg[v][u] = 0.f; }
    if (on[v]) {
      ldv<T>(static_cast<const T*>(p.bsf) + rpix * C + c, bs[v]);
      ldv<T>(dga + rpix * C + c, gg[v]);
#pragma unroll
      for (int u = 0; u < V; ++u) gg[v][u] = __fdiv_rn(gg[v][u], Lf);
    }
  }
  // One level pixel: gate / bsf gradients from d out, then d x = routed + d out
  // (the order of arfe_fpn_gather_backward_acc).  routed(v, u) = the gather's gradient.
  auto pixel = [&](int l, size_t pix, auto routed) {
    const TD* __restrict__ din = static_cast<const TD*>(p.feats[l]);
    T* __restrict__ dx = static_cast<T*>(p.outs[l]);
    const float a1 = ldf(static_cast<const T*>(p.g1[l]) + pix), a2 = ldf(static_cast<const T*>(p.g2[l]) + pix);
    const float t1 = tanhf(fmaxf(a1, 0.f)), t2 = tanhf(fmaxf(a2, 0.f));
    const float gate = t1 + t2;
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!on[v]) continue;
      float f[V];
      ldv_as_f32<TD, V>(din + pix * C + (v * 32 + lane) * V, f);
#pragma unroll
      for (int u = 0; u < V; ++u) {
        sum = fmaf(f[u], bs[v][u], sum);
        db[v][u] = fmaf(f[u], gate, db[v][u]);
        f[u] = routed(v, u) + f[u];
      }
      stv<T>(dx + pix * C + (v * 32 + lane) * V, f);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) {
      p.dg1[l][pix] = a1 > 0.f ? sum * (1.f - t1 * t1) : 0.f;
      p.dg2[l][pix] = a2 > 0.f ? sum * (1.f - t2 * t2) : 0.f;
    }
  };
  // ---- pooled levels: the s x s window of this refine pixel, argmax cell gets g / L ----
  for (int l = 0; l < p.refine_level; ++l) {
    const int H = p.H[l], W = p.W[l], s = ex.s[l];
    unsigned arg[NV][V / 4];
    const uint8_t* a = p.argmax + ((((size_t)l * p.B + b) * Hr + Y) * Wr + X) * C;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int u = 0; u < V / 4; ++u)
        arg[v][u] = on[v] ? __ldg(reinterpret_cast<const unsigned*>(a + (v * 32 + lane) * V) + u) : 0u;
    for (int dy = 0; dy < s; ++dy)
      for (int dxx = 0; dxx < s; ++dxx) {
        const unsigned pos = (unsigned)(dy * s + dxx);
        pixel(l, ((size_t)b * H + (size_t)s * Y + dy) * W + (size_t)s * X + dxx, [&](int v, int u) {
          return (((arg[v][u >> 2] >> (8 * (u & 3))) & 255u) == pos) ? gg[v][u] : 0.f;
        });
      }
  }
  // ---- the refine level itself ----
  pixel(p.refine_level, rpix, [&](int v, int u) { return gg[v][u]; });
  // ---- upsampled levels: the level pixel (at most one) whose nearest destination this is;
  // its gather gradient = sum of d(gathered) over the refine pixels that read it, / L ----
  for (int l = p.refine_level + 1; l < p.L; ++l) {
    const int H = p.H[l], W = p.W[l];
    int ya, yb, xa, xb;
    dst_range(Y, Hr, H, ya, yb);
    dst_range(X, Wr, W, xa, xb);
    for (int y = ya; y < yb; ++y)
      for (int x = xa; x < xb; ++x) {
        const float sy = (float)Hr / (float)H, sx = (float)Wr / (float)W;
        int Ya = (int)floorf((float)y * sy) - 1; if (Ya < 0) Ya = 0;
        int Yb = (int)ceilf((float)(y + 1) * sy) + 1; if (Yb > Hr) Yb = Hr;
        int Xa = (int)floorf((float)x * sx) - 1; if (Xa < 0) Xa = 0;
        int Xb = (int)ceilf((float)(x + 1) * sx) + 1; if (Xb > Wr) Xb = Wr;
        while (Ya < Yb && nearest_src(Ya, H, Hr) != y) ++Ya;
        while (Yb > Ya && nearest_src(Yb - 1, H, Hr) != y) --Yb;
        while (Xa < Xb && nearest_src(Xa, W, Wr) != x) ++Xa;
        while (Xb > Xa && nearest_src(Xb - 1, W, Wr) != x) --Xb;
        float r[NV][V];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int u = 0; u < V; ++u) r[v][u] = 0.f;
        for (int Y2 = Ya; Y2 < Yb; ++Y2)
          for (int X2 = Xa; X2 < Xb; ++X2)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              if (!on[v]) continue;
              float t[V];
              ldv<T>(dga + (((size_t)b * Hr + Y2) * Wr + X2) * C + (v * 32 + lane) * V, t);
#pragma unroll
              for (int u = 0; u < V; ++u) r[v][u] += t[u];
            }
        pixel(l, ((size_t)b * H + y) * W + x, [&](int v, int u) { return __fdiv_rn(r[v][u], Lf); });
      }
  }
  float* __restrict__ o = p.dbsf + rpix * C;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (!on[v]) continue;
    const int c = (v * 32 + lane) * V;
#pragma unroll
    for (int u = 0; u < V; u += 4)
      *reinterpret_cast<float4*>(o + c + u) = make_float4(db[v][u], db[v][u + 1], db[v][u + 2], db[v][u + 3]);
  }
}

// ---------------------------------------- fused AR-FPN backward, TMA-staged
// Same arithmetic as fpn_bwd_fused_cl, other data movement.  The register kernel keeps
// one level pixel (1 KB) in flight per warp and pays a DRAM latency per pixel: it runs at
// half the copy bandwidth however it is tuned.  In NHWC
// everything a refine pixel needs from d out is a handful of CONTIGUOUS pieces -- the s
// rows of its s x s window on every pooled level (s px x C channels each), its own pixel,
// the one pixel of each upsampled level that reads it -- so the whole footprint (23 KB at
// C = 256) is fetched with one bulk async copy (TMA) per piece into a shared-memory
// buffer, signalled by one mbarrier, and consumed from shared memory: 23 KB in flight per
// buffer instead of 1 KB per warp, no registers held by loads.
//   * a PAIR of warps shares a buffer, each warp taking half of the channels: sixteen
//     warps per SM hide each other's shared-memory and arithmetic latencies (eight
//     single warps with a buffer each issued one instruction every ~8 cycles);
//   * lane i of the pair's first warp owns level pixel i of the footprint for the scalar
//     work (gate maps in, d gate out): those loads and stores are parallel, not per pixel;
//   * fp32: d x = d out + gather gradient is formed IN the buffer -- the gather routes
//     g / L to one cell per channel and pooled level, i.e. one read-modify-write per
//     channel instead of a compare + select + add per channel and pixel -- and the rows
//     leave by bulk stores (TMA) as they came; bf16 d x is converted and stored from registers;
//   * d x of the levels above the refine level (a sum over the refine pixels reading one
//     level pixel) is finished by gather_bwd_up_cl with d out as the addend (2 % of the pyramid).
constexpr int kFusedPairs = 8;
constexpr int kFusedWarps = 2 * kFusedPairs;
constexpr int kFusedMaxJobs = 32;   // bulk copies per footprint (one per lane)
constexpr int kFusedMaxPix = 32;    // level pixels per footprint (one per lane)
constexpr int kFusedMaxPooled = 3;  // levels below the refine level

struct FusedGeom {
  int s[kMaxLevels];        // pooled levels: window side; 0 above the refine level
  int buf_bytes;            // footprint bytes, rounded up to 128
  int tab_bytes;            // upsampled-level lookup tables, rounded up to 128
  int items;                // B * Hr * Wr
};

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void pair_sync(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
}

// NV = 128-bit (fp32: 4 channels) / 8-channel (bf16) vectors per lane of ONE warp; the pair
// covers 2 * NV * 32 * V channels.
template <typename T, int NV>
__global__ void __launch_bounds__(kFusedWarps * 32, 1)
fpn_bwd_fused_tma(const FpnParams p, const FusedGeom geo) {
  constexpr int V = Vec<T>::n;
  constexpr bool kStoreTma = sizeof(T) == 4;  // d x has the staged element type
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, half = warp & 1;
  const int Hr = p.Hr, Wr = p.Wr, C = p.C, R = p.refine_level;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + pair;
  float* psum = reinterpret_cast<float*>(smem_raw + 128) + pair * 32;   // second warp's channel sums
  short* tab = reinterpret_cast<short*>(smem_raw + 128 + kFusedPairs * 128);  // level R+1+j: [Hr] rows, [Wr] columns
  unsigned char* buf = smem_raw + 128 + kFusedPairs * 128 + geo.tab_bytes + (size_t)pair * geo.buf_bytes;
  if (half == 0 && lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // which row / column of an upsampled level reads refine row Y / column X (-1: none)
  for (int e = threadIdx.x; e < (p.L - 1 - R) * (Hr + Wr); e += blockDim.x) {
    const int l = R + 1 + e / (Hr + Wr), i = e % (Hr + Wr);
    int a, bnd;
    if (i < Hr) dst_range(i, Hr, p.H[l], a, bnd);
    else dst_range(i - Hr, Wr, p.W[l], a, bnd);
    tab[e] = (short)(bnd > a ? a : -1);
  }
  __syncthreads();
  const uint32_t pixB = (uint32_t)C * 4u;  // d out is fp32
  const float Lf = (float)p.L;
  bool on[NV];
  int ch[NV];  // first channel of this lane's v-th vector
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    ch[v] = ((half * NV + v) * 32 + lane) * V;
    on[v] = ch[v] < C;
  }
  uint32_t phase = 0;
  for (int item = blockIdx.x * kFusedPairs + pair; item < geo.items; item += gridDim.x * kFusedPairs) {
    const int X = item % Wr;
    const int Y = (item / Wr) % Hr;
    const int b = item / (Wr * Hr);
    const size_t rpix = ((size_t)b * Hr + Y) * Wr + X;
    // ---- footprint: copy jobs (lane j of the first warp moves job j, in and out) and level
    // pixels (lane i owns pixel i); both warps compute it, it is a few dozen instructions
    const float* src = nullptr;
    float* gdst = nullptr;      // where this lane's piece goes as d x (fp32, levels <= R)
    uint32_t dst = 0, bytes = 0, off = 0, off_r = 0;
    int job = 0, npix = 0;
    int my_l = -1;
    size_t my_pix = 0;
    uint32_t loff[kMaxLevels];  // buffer offset of level l's pixels
    bool have[kMaxLevels];
    size_t lpix[kMaxLevels];    // global pixel index of the piece's first pixel
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      loff[l] = off;
      have[l] = false;
      lpix[l] = 0;
      if (l >= p.L) continue;
      const int H = p.H[l], W = p.W[l];
      int ya, xa, n;  // piece = n rows x n pixels from (ya, xa)
      if (l < R) { n = geo.s[l]; ya = n * Y; xa = n * X; }
      else if (l == R) { n = 1; ya = Y; xa = X; }
      else {
        const short* t = tab + (l - R - 1) * (Hr + Wr);
        ya = t[Y]; xa = t[Hr + X]; n = (ya >= 0 && xa >= 0) ? 1 : 0;
      }
      if (n == 0) continue;
      have[l] = true;
      if (l == R) off_r = off;
      lpix[l] = ((size_t)b * H + ya) * W + xa;
      const uint32_t rowB = (uint32_t)n * pixB;
      if (half == 0 && lane >= job && lane < job + n) {
        const int r = lane - job;
        src = static_cast<const float*>(p.feats[l]) + (lpix[l] + (size_t)r * W) * C;
        if (kStoreTma && l <= R) gdst = static_cast<float*>(p.outs[l]) + (lpix[l] + (size_t)r * W) * C;
        dst = off + (uint32_t)r * rowB;
        bytes = rowB;
      }
      if (lane >= npix && lane < npix + n * n) {
        const int q = lane - npix;
        const int sh = 31 - __clz(n);
        const int qr = (n & (n - 1)) == 0 ? q >> sh : q / n;  // window sides are 1, 2, 4 in practice
        my_l = l;
        my_pix = lpix[l] + (size_t)qr * W + (q - qr * n);
      }
      job += n;
      npix += n * n;
      off += (uint32_t)n * rowB;
    }
    float a1 = 0.f, a2 = 0.f;  // gate values of this lane's level pixel: in flight first
    if (my_l >= 0) {
      a1 = ldf(static_cast<const T*>(p.g1[my_l]) + my_pix);
      a2 = ldf(static_cast<const T*>(p.g2[my_l]) + my_pix);
    }
    // ---- first warp: issue the loads once its bulk stores of the previous item have read the buffer
    // (the second warp's last access to the buffer precedes the pair barrier of the previous item)
    if (half == 0) {
      if constexpr (kStoreTma) {
        if (gdst) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, off);
      __syncwarp();
      if (bytes) bulk_g2s(buf + dst, src, bytes, bar);
    }
    // ---- plain loads, in flight with the copies: gate values of this lane's level pixel,
    // bsf / d(gathered) / argmax bytes of this lane's channels
    float bs[NV][V], db[NV][V], gg[NV][V];
    unsigned arg[kFusedMaxPooled][NV][V / 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int u = 0; u < V; ++u) { bs[v][u] = 0.f; db[v][u] = 0.f; gg[v][u] = 0.f; }
#pragma unroll
      for (int l = 0; l < kFusedMaxPooled; ++l)
#pragma unroll
        for (int u = 0; u < V / 4; ++u) arg[l][v][u] = 0u;
      if (on[v]) {
        ldv<T>(static_cast<const T*>(p.bsf) + rpix * C + ch[v], bs[v]);
        ldv<T>(static_cast<const T*>(p.gathered) + rpix * C + ch[v], gg[v]);
#pragma unroll
        for (int l = 0; l < kFusedMaxPooled; ++l)
          if (l < R) {
            const uint8_t* a = p.argmax + (((size_t)l * p.B + b) * Hr * Wr + (size_t)Y * Wr + X) * C + ch[v];
#pragma unroll
            for (int u = 0; u < V / 4; ++u) arg[l][v][u] = __ldg(reinterpret_cast<const unsigned*>(a) + u);
          }
      }
    }
    float t1 = 0.f, t2 = 0.f;
    if (my_l >= 0) {
      t1 = tanhf(fmaxf(a1, 0.f));
      t2 = tanhf(fmaxf(a2, 0.f));
    }
    const float mygate = t1 + t2;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int u = 0; u < V; ++u) gg[v][u] = __fdiv_rn(gg[v][u], Lf);
    mbar_wait(bar, phase);
    phase ^= 1u;
    // ---- the level pixels, from shared memory
    int q = 0;            // running level-pixel index == owning lane
    float mysum = 0.f;    // this warp's channel sum of the pixel this lane owns
    // lane-local part of one level pixel: returns this lane's share of sum_c d out * bsf
    auto pixel = [&](int l, uint32_t boff, size_t pix, int qq, bool write_dx, auto routed) -> float {
      const float gate = __shfl_sync(0xffffffffu, mygate, qq);
      float sum[4] = {0.f, 0.f, 0.f, 0.f};  // four chains: the dot product is not one dependent FMA string
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!on[v]) continue;
        float f[V];
        const float* sp = reinterpret_cast<const float*>(buf + boff) + ch[v];
#pragma unroll
        for (int u = 0; u < V; u += 4) {
          const float4 t = *reinterpret_cast<const float4*>(sp + u);
          f[u] = t.x; f[u + 1] = t.y; f[u + 2] = t.z; f[u + 3] = t.w;
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
          sum[u & 3] = fmaf(f[u], bs[v][u], sum[u & 3]);
          db[v][u] = fmaf(f[u], gate, db[v][u]);
        }
        if constexpr (!kStoreTma) {
          if (write_dx) {
#pragma unroll
            for (int u = 0; u < V; ++u) f[u] = routed(v, u) + f[u];
            stv<T>(static_cast<T*>(p.outs[l]) + pix * C + ch[v], f);
          }
        }
      }
      return (sum[0] + sum[1]) + (sum[2] + sum[3]);
    };
    // warp totals of N per-lane values at once (N = 4: 7 shuffles instead of 20): after the
    // exchange steps lane L holds the total of value (L >> 3) & (N - 1); pixel q0 + i goes to lane q0 + i
    auto reduce_to_owners = [&](auto n_tag, const float (&val)[decltype(n_tag)::value], int q0) {
      constexpr int N = decltype(n_tag)::value;
      float x;
      if constexpr (N == 4) {
        const bool hi = (lane & 16) != 0, mid = (lane & 8) != 0;
        float k0 = hi ? val[2] : val[0], k1 = hi ? val[3] : val[1];
        k0 += __shfl_xor_sync(0xffffffffu, hi ? val[0] : val[2], 16);
        k1 += __shfl_xor_sync(0xffffffffu, hi ? val[1] : val[3], 16);
        x = (mid ? k1 : k0) + __shfl_xor_sync(0xffffffffu, mid ? k0 : k1, 8);
      } else if constexpr (N == 2) {
        const bool mid = (lane & 8) != 0;
        x = (mid ? val[1] : val[0]) + __shfl_xor_sync(0xffffffffu, mid ? val[0] : val[1], 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
      } else {
        x = val[0] + __shfl_xor_sync(0xffffffffu, val[0], 16);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
      }
      x += __shfl_xor_sync(0xffffffffu, x, 4);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      const int i = lane - q0;
      const float got = __shfl_sync(0xffffffffu, x, (i & (N - 1)) << 3);
      if (i >= 0 && i < N) mysum = got;
    };
    auto window = [&](auto s_tag, int l) {  // pooled level, window side known at compile time
      constexpr int S = decltype(s_tag)::value;
      const int W = p.W[l];
#pragma unroll
      for (int dy = 0; dy < S; ++dy) {
        float val[S];
#pragma unroll
        for (int dxx = 0; dxx < S; ++dxx) {
          const unsigned pos = (unsigned)(dy * S + dxx);
          val[dxx] = pixel(l, loff[l] + pos * pixB, lpix[l] + (size_t)dy * W + dxx, q + dxx, true, [&](int v, int u) {
            return (((arg[l < kFusedMaxPooled ? l : 0][v][u >> 2] >> (8 * (u & 3))) & 255u) == pos) ? gg[v][u] : 0.f;
          });
        }
        reduce_to_owners(std::integral_constant<int, S>{}, val, q);
        q += S;
      }
    };
    auto single = [&](int l, uint32_t boff, size_t pix, bool write_dx, auto routed) {
      float val[1] = {pixel(l, boff, pix, q, write_dx, routed)};
      reduce_to_owners(std::integral_constant<int, 1>{}, val, q);
      ++q;
    };
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      if (l >= p.L || !have[l]) continue;
      if (l < R) {
        const int s = geo.s[l];
        if (s == 4) window(std::integral_constant<int, 4>{}, l);
        else if (s == 2) window(std::integral_constant<int, 2>{}, l);
        else {
          const int W = p.W[l];
          for (int dy = 0; dy < s; ++dy)
            for (int dxx = 0; dxx < s; ++dxx) {
              const unsigned pos = (unsigned)(dy * s + dxx);
              single(l, loff[l] + pos * pixB, lpix[l] + (size_t)dy * W + dxx, true, [&](int v, int u) {
                return (((arg[l < kFusedMaxPooled ? l : 0][v][u >> 2] >> (8 * (u & 3))) & 255u) == pos) ? gg[v][u] : 0.f;
              });
            }
        }
      } else if (l == R) {
        single(l, loff[l], rpix, true, [&](int v, int u) { return gg[v][u]; });
      } else {
        single(l, loff[l], 0, false, [&](int, int) { return 0.f; });  // d x: gather_bwd_up_cl
      }
    }
    if (half == 1) psum[lane] = mysum;
    if constexpr (kStoreTma) {
      // d x in place: the gather's gradient goes to the argmax cell of every pooled window
      // (one read-modify-write per channel and level) and to the refine pixel itself.
      // Each lane touches only its own channels, which no other lane of the pair reads.
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!on[v]) continue;
        const uint32_t coff = (uint32_t)ch[v] * 4u;
#pragma unroll
        for (int l = 0; l < kFusedMaxPooled; ++l)
          if (l < R) {
#pragma unroll
            for (int u = 0; u < V; ++u) {
              const unsigned pos = (arg[l][v][u >> 2] >> (8 * (u & 3))) & 255u;
              float* cell = reinterpret_cast<float*>(buf + loff[l] + pos * pixB + coff) + u;
              *cell = gg[v][u] + *cell;
            }
          }
        float* cell = reinterpret_cast<float*>(buf + off_r + coff);
#pragma unroll
        for (int u = 0; u < V; ++u) cell[u] = gg[v][u] + cell[u];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pair_sync(pair);  // both halves of every pixel are final; the second warp's sums are visible
    if (half == 0) {
      if constexpr (kStoreTma) {
        if (gdst) {
          bulk_s2g(gdst, buf + dst, bytes);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (my_l >= 0) {
        const float tot = mysum + psum[lane];
        p.dg1[my_l][my_pix] = a1 > 0.f ? tot * (1.f - t1 * t1) : 0.f;
        p.dg2[my_l][my_pix] = a2 > 0.f ? tot * (1.f - t2 * t2) : 0.f;
      }
    }
    float* __restrict__ o = p.dbsf + rpix * C;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!on[v]) continue;
#pragma unroll
      for (int u = 0; u < V; u += 4)
        *reinterpret_cast<float4*>(o + ch[v] + u) = make_float4(db[v][u], db[v][u + 1], db[v][u + 2], db[v][u + 3]);
    }
    __syncwarp();
  }
  if constexpr (kStoreTma) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Pooling ratio of each level below the refine level when it is an exact integer.
inline Exact exact_ratios(const FpnParams& p) {
  Exact ex;
  for (int l = 0; l < kMaxLevels; ++l) ex.s[l] = 0;
  for (int l = 0; l < p.refine_level; ++l)
    if (p.H[l] % p.Hr == 0 && p.W[l] % p.Wr == 0 && p.H[l] / p.Hr == p.W[l] / p.Wr && p.H[l] / p.Hr <= 15)
      ex.s[l] = p.H[l] / p.Hr;
  return ex;
}

inline unsigned blocks_for(size_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }
inline bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// True when the vector kernels can take this problem.
bool fpn_cl_ok(const FpnParams& p, int dtype, bool need_feats, bool need_outs) {
  const int V = dtype == 0 ? 4 : 8;
  if (p.C % V) return false;
  for (int l = 0; l < p.L; ++l) {
    if (need_feats && !a16(p.feats[l])) return false;
    if (need_outs && !a16(p.outs[l])) return false;
  }
  return true;
}

cudaError_t launch_fpn_gather_forward_cl(const FpnParams& p, int dtype, cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const size_t total = (size_t)p.B * p.Hr * p.Wr * (p.C / V);
  if (total == 0) return cudaSuccess;
  const Exact ex = exact_ratios(p);
  if (dtype == 0) gather_fwd_cl<float><<<blocks_for(total, kThreads), kThreads, 0, stream>>>(p, ex);
  else gather_fwd_cl<__nv_bfloat16><<<blocks_for(total, kThreads), kThreads, 0, stream>>>(p, ex);
  return cudaGetLastError();
}

// Returns in *mask the levels NOT written here (generic kernel finishes them).
cudaError_t launch_fpn_gather_backward_cl(const FpnParams& p, int dtype, unsigned* mask,
                                          cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const Exact ex = exact_ratios(p);
  for (int l = 0; l < p.L; ++l)
    if (p.addend[l] && !a16(p.addend[l])) return cudaErrorInvalidValue;
  unsigned m = (1u << p.L) - 1u;
  for (int l = 0; l < p.refine_level; ++l)
    if (ex.s[l]) m &= ~(1u << l);
  m &= ~(1u << p.refine_level);
  const size_t total = (size_t)p.B * p.Hr * p.Wr * (p.C / V);
  if (total == 0) { *mask = 0; return cudaSuccess; }
  if (dtype == 0) gather_bwd_cl<float><<<blocks_for(total, kThreads), kThreads, 0, stream>>>(p, ex);
  else gather_bwd_cl<__nv_bfloat16><<<blocks_for(total, kThreads), kThreads, 0, stream>>>(p, ex);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // levels above the refine level: vector kernel
  UpLevels ul;
  size_t st = 0;
  int n = 0;
  for (int l = p.refine_level + 1; l < p.L; ++l) {
    ul.start[n] = st;
    ul.level[n] = l;
    st += (size_t)p.B * p.H[l] * p.W[l] * (p.C / V);
    m &= ~(1u << l);
    ++n;
  }
  ul.n = n;
  for (int j = n; j <= kMaxLevels; ++j) ul.start[j] = st;
  for (int j = n; j < kMaxLevels; ++j) ul.level[j] = 0;
  *mask = m;
  if (st) {
    if (dtype == 0) gather_bwd_up_cl<float><<<blocks_for(st, kThreads), kThreads, 0, stream>>>(p, ul);
    else gather_bwd_up_cl<__nv_bfloat16><<<blocks_for(st, kThreads), kThreads, 0, stream>>>(p, ul);
  }
  return cudaGetLastError();
}

template <bool kBackward>
static cudaError_t launch_apply_cl(const FpnParams& p, int dtype, cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const size_t warps = (size_t)p.B * p.Hr * p.Wr;
  if (warps == 0) return cudaSuccess;
  const unsigned grid = blocks_for(warps, kThreads / 32);
  const int nv = (p.C + 32 * V - 1) / (32 * V);
  if (nv > 4) return cudaErrorInvalidValue;
  // resident CTAs per SM the compiler must allow: the backward is latency-bound, its
  // speed follows occupancy (80 registers: 76 us, 64: 62 us); forward 4 / 2
  static const int occ_env = ARFE_KNOB_ENV("ARFE_APPLY_OCC", 0);
  const int occ = kBackward ? (occ_env ? occ_env : 4) : 0;
#define ARFE_APPLY(TT, NV)                                                                         \
  do {                                                                                             \
    if (!kBackward) apply_cl<TT, NV, kBackward, (NV * Vec<TT>::n <= 8 ? 4 : 2)><<<grid, kThreads, 0, stream>>>(p); \
    else if (occ == 5) apply_cl<TT, NV, kBackward, 5><<<grid, kThreads, 0, stream>>>(p);           \
    else if (occ == 6) apply_cl<TT, NV, kBackward, 6><<<grid, kThreads, 0, stream>>>(p);           \
    else apply_cl<TT, NV, kBackward, 4><<<grid, kThreads, 0, stream>>>(p);                         \
  } while (0)
  if (dtype == 0) {
    switch (nv) { case 1: ARFE_APPLY(float, 1); break; case 2: ARFE_APPLY(float, 2); break;
                  case 3: ARFE_APPLY(float, 3); break; default: ARFE_APPLY(float, 4); break; }
  } else {
    switch (nv) { case 1: ARFE_APPLY(__nv_bfloat16, 1); break; case 2: ARFE_APPLY(__nv_bfloat16, 2); break;
                  case 3: ARFE_APPLY(__nv_bfloat16, 3); break; default: ARFE_APPLY(__nv_bfloat16, 4); break; }
  }
#undef ARFE_APPLY
  return cudaGetLastError();
}

cudaError_t launch_fpn_apply_forward_cl(const FpnParams& p, int dtype, cudaStream_t stream) {
  return launch_apply_cl<false>(p, dtype, stream);
}
cudaError_t launch_fpn_apply_backward_cl(const FpnParams& p, int dtype, cudaStream_t stream) {
  return launch_apply_cl<true>(p, dtype, stream);
}

// dout_f32: the incoming gradients are fp32 whatever the feature dtype is.
// cudaErrorNotSupported: not a case of the fused kernel (the caller uses the two calls).
cudaError_t launch_fpn_backward_fused_cl(const FpnParams& p, int dtype, int dout_f32, cudaStream_t stream) {
  const int V = dtype == 0 ? 4 : 8;
  const size_t warps = (size_t)p.B * p.Hr * p.Wr;
  if (warps == 0) return cudaSuccess;
  const int nv = (p.C + 32 * V - 1) / (32 * V);
  if (p.C % V || nv > 2) return cudaErrorNotSupported;
  const Exact ex = exact_ratios(p);
  for (int l = 0; l < p.L; ++l) {
    if (l < p.refine_level && (ex.s[l] == 0 || ex.s[l] > 15)) return cudaErrorNotSupported;
    if (!a16(p.feats[l]) || !a16(p.outs[l])) return cudaErrorNotSupported;
  }
  if (!a16(p.bsf) || !a16(p.gathered) || !a16(p.dbsf) || (reinterpret_cast<uintptr_t>(p.argmax) & 3u))
    return cudaErrorNotSupported;
  // ---- TMA-staged kernel: fp32 d out, footprint of a refine pixel within one lane-set ----
  if ((dtype == 0 || dout_f32) && p.C % 16 == 0) {
    FusedGeom geo;
    int jobs = 0, pix = 0;
    size_t bytes = 0;
    for (int l = 0; l < kMaxLevels; ++l) geo.s[l] = 0;
    for (int l = 0; l < p.L; ++l) {
      if (l < p.refine_level) { geo.s[l] = ex.s[l]; jobs += ex.s[l]; pix += ex.s[l] * ex.s[l]; bytes += (size_t)ex.s[l] * ex.s[l]; }
      else if (l > p.refine_level) {
        // upsampled level (a map no larger than the refine map): nearest_src(y: Hr <- H) is strictly
        // increasing in y, so at most one level pixel reads a given refine pixel
        if (p.H[l] > p.Hr || p.W[l] > p.Wr) { jobs = 1 << 20; break; }
        jobs += 1; pix += 1; bytes += 1;
      } else { jobs += 1; pix += 1; bytes += 1; }
    }
    if (p.refine_level > kFusedMaxPooled) jobs = 1 << 20;
    bytes = bytes * p.C * 4;
    bytes = (bytes + 127) / 128 * 128;
    size_t tab = (size_t)(p.L - 1 - p.refine_level) * (p.Hr + p.Wr) * 2;
    tab = (tab + 127) / 128 * 128;
    const size_t smem = 128 + kFusedPairs * 128 + tab + bytes * kFusedPairs;
    const int nvw = (nv + 1) / 2;  // vectors per lane of one warp of a pair
    if (jobs <= kFusedMaxJobs && pix <= kFusedMaxPix && smem <= 220 * 1024 && warps < (1u << 30) && nvw == 1) {
      geo.buf_bytes = (int)bytes;
      geo.tab_bytes = (int)tab;
      geo.items = (int)warps;
      int grid = sm_count();
      const int need = (int)((warps + kFusedPairs - 1) / kFusedPairs);
      if (grid > need) grid = need;
      cudaError_t e;
#define ARFE_FUSED_TMA(TT)                                                                            \
  do {                                                                                                \
    if ((e = cudaFuncSetAttribute(fpn_bwd_fused_tma<TT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e; \
    fpn_bwd_fused_tma<TT, 1><<<grid, kFusedWarps * 32, smem, stream>>>(p, geo);                       \
  } while (0)
      if (dtype == 0) ARFE_FUSED_TMA(float);
      else ARFE_FUSED_TMA(__nv_bfloat16);
#undef ARFE_FUSED_TMA
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      // d x of the levels above the refine level: gather gradient + d out
      FpnParams q = p;
      UpLevels ul;
      size_t st = 0;
      int n = 0;
      for (int l = 0; l < kMaxLevels; ++l) q.addend[l] = nullptr;
      for (int l = p.refine_level + 1; l < p.L; ++l) {
        ul.start[n] = st;
        ul.level[n] = l;
        st += (size_t)p.B * p.H[l] * p.W[l] * (p.C / V);
        q.addend[l] = static_cast<const float*>(p.feats[l]);
        ++n;
      }
      ul.n = n;
      for (int j = n; j <= kMaxLevels; ++j) ul.start[j] = st;
      for (int j = n; j < kMaxLevels; ++j) ul.level[j] = 0;
      if (st) {
        if (dtype == 0) gather_bwd_up_cl<float><<<blocks_for(st, kThreads), kThreads, 0, stream>>>(q, ul);
        else gather_bwd_up_cl<__nv_bfloat16><<<blocks_for(st, kThreads), kThreads, 0, stream>>>(q, ul);
      }
      return cudaGetLastError();
    }
  }
  const unsigned grid = blocks_for(warps, kThreads / 32);
#define ARFE_FUSED(TT, TD, NV) fpn_bwd_fused_cl<TT, TD, NV><<<grid, kThreads, 0, stream>>>(p, ex)
  if (dtype == 0) {
    if (nv == 1) ARFE_FUSED(float, float, 1); else ARFE_FUSED(float, float, 2);
  } else if (dout_f32) {
    if (nv == 1) ARFE_FUSED(__nv_bfloat16, float, 1); else ARFE_FUSED(__nv_bfloat16, float, 2);
  } else {
    if (nv == 1) ARFE_FUSED(__nv_bfloat16, __nv_bfloat16, 1); else ARFE_FUSED(__nv_bfloat16, __nv_bfloat16, 2);
  }
#undef ARFE_FUSED
  return cudaGetLastError();
}

}  // namespace arfe
