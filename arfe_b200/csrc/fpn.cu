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
#include "fpn_common.cuh"

namespace arfe {
using namespace fpn;
namespace {

// argmax buffer element of pooled level l: laid out like the tensors
// (NCHW: [l][B][C][Hr][Wr], channels-last: [l][B][Hr][Wr][C]).
template <bool kNHWC>
__device__ __forceinline__ size_t amax_at(int l, int B, int b, int c, int Y, int X, int C, int Hr, int Wr) {
  return kNHWC ? ((((size_t)l * B + b) * Hr + Y) * Wr + X) * C + c
               : (((size_t)l * B + b) * C + c) * Hr * Wr + (size_t)Y * Wr + X;
}

// ---------------------------------------------------------------- gather fwd
// Window maximum with the reference's tie/NaN rule, element order row-major.
__device__ __forceinline__ void max_step(float t, int pos, float& best, int& arg) {
  if (t > best || t != t) { best = t; arg = pos; }
}

// s x s window of an NCHW fp32 plane read with 128/64-bit loads (exact ratio).
template <int S>
__device__ __forceinline__ void window_max_vec(const float* __restrict__ row0, int W,
                                               float& best, int& arg) {
#pragma unroll
  for (int r = 0; r < S; ++r) {
    const float* q = row0 + (size_t)r * W;
    if constexpr (S == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(q));
      max_step(v.x, r * 4 + 0, best, arg); max_step(v.y, r * 4 + 1, best, arg);
      max_step(v.z, r * 4 + 2, best, arg); max_step(v.w, r * 4 + 3, best, arg);
    } else {
      const float2 v = __ldg(reinterpret_cast<const float2*>(q));
      max_step(v.x, r * 2 + 0, best, arg); max_step(v.y, r * 2 + 1, best, arg);
    }
  }
}

// `vec[l]` = pooling ratio (2 or 4) when level l < refine has an exact integer
// ratio, an NCHW fp32 layout and 16-byte aligned rows; 0 selects generic loops.
struct VecFlags { int s[kMaxLevels]; };

template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads)
gather_fwd(const FpnParams p, const VecFlags vf) {
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
      float best = -CUDART_INF_F;
      int arg = 0;
      bool done = false;
      if constexpr (sizeof(T) == 4 && !kNHWC) {
        const int s = vf.s[l];
        if (s == 4) {
          window_max_vec<4>(reinterpret_cast<const float*>(f) + (((size_t)b * C + c) * H + 4 * Y) * W + 4 * X, W, best, arg);
          done = true;
        } else if (s == 2) {
          window_max_vec<2>(reinterpret_cast<const float*>(f) + (((size_t)b * C + c) * H + 2 * Y) * W + 2 * X, W, best, arg);
          done = true;
        }
      }
      if (!done) {
        const int y0 = pool_start(Y, H, Hr), y1 = pool_end(Y, H, Hr);
        const int x0 = pool_start(X, W, Wr), x1 = pool_end(X, W, Wr);
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x)
            max_step(ldf(f + at<kNHWC>(b, c, y, x, C, H, W)), (y - y0) * (x1 - x0) + (x - x0), best, arg);
      }
      v = best;
      if (p.argmax)
        p.argmax[amax_at<kNHWC>(l, p.B, b, c, Y, X, C, Hr, Wr)] = (uint8_t)arg;
    } else {
      v = ldf(f + at<kNHWC>(b, c, nearest_src(Y, H, Hr), nearest_src(X, W, Wr), C, H, W));
    }
    acc = __fadd_rn(acc, v);
  }
  stf(static_cast<T*>(p.gathered) + i, __fdiv_rn(acc, (float)p.L));
}

// ---------------------------------------------------------------- gather bwd
// Generic: one thread per element of the gradient of each level listed in
// `lo.level[]` (all fully written).
struct LevelOffsets {
  size_t start[kMaxLevels + 1];
  int level[kMaxLevels];
  int n;
};

template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads)
gather_bwd(const FpnParams p, const LevelOffsets lo) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lo.start[lo.n]) return;
  int j = 0;
  while (j + 1 < lo.n && i >= lo.start[j + 1]) ++j;
  const int l = lo.level[j];
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  int b, c, y, x;
  decode<kNHWC>(i - lo.start[j], C, H, W, b, c, y, x);
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
        const int arg = p.argmax[amax_at<kNHWC>(l, p.B, b, c, Y, X, C, Hr, Wr)];
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
  g = __fdiv_rn(g, (float)p.L);
  if (p.addend[l]) g += __ldg(p.addend[l] + (i - lo.start[j]));
  stf(static_cast<T*>(p.outs[l]) + (i - lo.start[j]), g);
}

// Fast (NCHW fp32, exact integer ratios): one thread per refine element routes
// g/L into the s x s block of every pooled level with 128/64-bit stores and
// writes the refine level itself.
__global__ void __launch_bounds__(kThreads)
gather_bwd_fast(const FpnParams p, const VecFlags vf) {
  const int Hr = p.Hr, Wr = p.Wr, C = p.C;
  const size_t total = (size_t)p.B * C * Hr * Wr;
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= total) return;
  int b, c, Y, X;
  decode<false>(i, C, Hr, Wr, b, c, Y, X);
  const float g = __fdiv_rn(__ldg(static_cast<const float*>(p.gathered) + i), (float)p.L);
  for (int l = 0; l < p.refine_level; ++l) {
    const int s = vf.s[l];
    if (s == 0) continue;
    const int W = p.W[l], H = p.H[l];
    const int arg = p.argmax[(((size_t)l * p.B + b) * C + c) * Hr * Wr + (size_t)Y * Wr + X];
    const size_t at0 = (((size_t)b * C + c) * H + (size_t)s * Y) * W + (size_t)s * X;
    float* __restrict__ o = static_cast<float*>(p.outs[l]) + at0;
    const float* __restrict__ add = p.addend[l] ? p.addend[l] + at0 : nullptr;
    if (s == 4) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = arg - 4 * r;
        float4 v = make_float4(k == 0 ? g : 0.f, k == 1 ? g : 0.f, k == 2 ? g : 0.f, k == 3 ? g : 0.f);
        if (add) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(add + (size_t)r * W));
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        *reinterpret_cast<float4*>(o + (size_t)r * W) = v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int k = arg - 2 * r;
        float2 v = make_float2(k == 0 ? g : 0.f, k == 1 ? g : 0.f);
        if (add) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(add + (size_t)r * W));
          v.x += t.x; v.y += t.y;
        }
        *reinterpret_cast<float2*>(o + (size_t)r * W) = v;
      }
    }
  }
  static_cast<float*>(p.outs[p.refine_level])[i] = p.addend[p.refine_level] ? g + __ldg(p.addend[p.refine_level] + i) : g;
}

// ----------------------------------------------------------------- apply fwd

// NCHW: thread = one pixel of one level, loops over a group of channels so
// the two tanh are amortised; consecutive threads = consecutive x (coalesced).
template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_fwd_nchw(const FpnParams p, const LevelOffsets lo /* pixel offsets */, int cg) {
  const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= lo.start[lo.n]) return;
  int j = 0;
  while (j + 1 < lo.n && i >= lo.start[j + 1]) ++j;
  const int l = lo.level[j];
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  size_t r = i - lo.start[j];
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
  if (i >= lo.start[lo.n]) return;
  int j = 0;
  while (j + 1 < lo.n && i >= lo.start[j + 1]) ++j;
  const int l = lo.level[j];
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  int b, c, y, x;
  decode<true>(i - lo.start[j], C, H, W, b, c, y, x);
  const size_t pix = ((size_t)b * H + y) * W + x;
  const float gate = gate_value(ldf(static_cast<const T*>(p.g1[l]) + pix),
                                ldf(static_cast<const T*>(p.g2[l]) + pix));
  const int ny = nearest_src(y, Hr, H), nx = nearest_src(x, Wr, W);
  const float bv = ldf(static_cast<const T*>(p.bsf) + at<true>(b, c, ny, nx, C, Hr, Wr));
  const size_t e = i - lo.start[j];
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
  if (i >= lo.start[lo.n]) return;
  int j = 0;
  while (j + 1 < lo.n && i >= lo.start[j + 1]) ++j;
  const int l = lo.level[j];
  const int H = p.H[l], W = p.W[l], C = p.C, Hr = p.Hr, Wr = p.Wr;
  size_t r = i - lo.start[j];
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

inline LevelOffsets offsets(const FpnParams& p, bool per_pixel, unsigned level_mask = 0xffffffffu) {
  LevelOffsets lo;
  size_t s = 0;
  int n = 0;
  for (int l = 0; l < p.L; ++l) {
    if (!(level_mask & (1u << l))) continue;
    lo.start[n] = s;
    lo.level[n] = l;
    s += (size_t)p.B * (per_pixel ? 1 : p.C) * p.H[l] * p.W[l];
    ++n;
  }
  lo.n = n;
  for (int j = n; j <= kMaxLevels; ++j) lo.start[j] = s;
  for (int j = n; j < kMaxLevels; ++j) lo.level[j] = 0;
  return lo;
}

// Levels below refine whose pooling ratio is an exact 2 or 4 with 16-byte
// aligned rows (NCHW fp32 only) take the vector paths.
inline VecFlags vec_flags(const FpnParams& p, int dtype, int layout, bool for_outs) {
  VecFlags vf;
  for (int l = 0; l < kMaxLevels; ++l) vf.s[l] = 0;
  if (dtype != 0 || layout != 0) return vf;
  for (int l = 0; l < p.refine_level; ++l) {
    for (int s = 2; s <= 4; s *= 2) {
      const void* base = for_outs ? p.outs[l] : p.feats[l];
      const size_t align = s == 4 ? 16 : 8;
      if (p.H[l] == s * p.Hr && p.W[l] == s * p.Wr &&
          (reinterpret_cast<uintptr_t>(base) % align) == 0)
        vf.s[l] = s;
    }
  }
  return vf;
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
  if (layout == 1 && fpn_cl_ok(p, dtype, true, false) &&
      (reinterpret_cast<uintptr_t>(p.gathered) & 15u) == 0 &&
      (p.argmax == nullptr || (reinterpret_cast<uintptr_t>(p.argmax) & 3u) == 0))
    return launch_fpn_gather_forward_cl(p, dtype, stream);
  const VecFlags vf = vec_flags(p, dtype, layout, false);
  ARFE_DISPATCH(gather_fwd, blocks_for(total), p, vf);
  return cudaGetLastError();
}

cudaError_t launch_fpn_gather_backward(const FpnParams& p, int dtype, int layout,
                                       cudaStream_t stream) {
  if (layout == 1 && fpn_cl_ok(p, dtype, false, true) &&
      (reinterpret_cast<uintptr_t>(p.gathered) & 15u) == 0 &&
      (p.argmax == nullptr || (reinterpret_cast<uintptr_t>(p.argmax) & 3u) == 0)) {
    unsigned rest = 0;
    cudaError_t e = launch_fpn_gather_backward_cl(p, dtype, &rest, stream);
    if (e != cudaSuccess) return e;
    const LevelOffsets lo = offsets(p, false, rest);
    if (lo.start[lo.n] == 0) return cudaSuccess;
    ARFE_DISPATCH(gather_bwd, blocks_for(lo.start[lo.n]), p, lo);
    return cudaGetLastError();
  }
  const VecFlags vf = vec_flags(p, dtype, layout, true);
  unsigned mask = (1u << p.L) - 1u;
  bool any_fast = false;
  for (int l = 0; l < p.refine_level; ++l)
    if (vf.s[l]) { mask &= ~(1u << l); any_fast = true; }
  if (any_fast) {
    mask &= ~(1u << p.refine_level);
    const size_t total = (size_t)p.B * p.C * p.Hr * p.Wr;
    if (total) gather_bwd_fast<<<blocks_for(total), kThreads, 0, stream>>>(p, vf);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const LevelOffsets lo = offsets(p, false, mask);
  if (lo.start[lo.n] == 0) return cudaSuccess;
  ARFE_DISPATCH(gather_bwd, blocks_for(lo.start[lo.n]), p, lo);
  return cudaGetLastError();
}

static bool apply_cl_ok(const FpnParams& p, int dtype) {
  const int V = dtype == 0 ? 4 : 8;
  return p.C <= 4 * 32 * V && (reinterpret_cast<uintptr_t>(p.bsf) & 15u) == 0;
}

cudaError_t launch_fpn_apply_forward(const FpnParams& p, int dtype, int layout,
                                     cudaStream_t stream) {
  if (layout == 1 && fpn_cl_ok(p, dtype, true, true) && apply_cl_ok(p, dtype))
    return launch_fpn_apply_forward_cl(p, dtype, stream);
  if (layout == 0) {
    const LevelOffsets lo = offsets(p, true);
    if (lo.start[lo.n] == 0) return cudaSuccess;
    const int cg = p.C >= 32 ? 32 : p.C;
    dim3 grid(blocks_for(lo.start[lo.n]), (p.C + cg - 1) / cg);
    if (dtype == 0) apply_fwd_nchw<float><<<grid, kThreads, 0, stream>>>(p, lo, cg);
    else apply_fwd_nchw<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(p, lo, cg);
  } else {
    const LevelOffsets lo = offsets(p, false);
    if (lo.start[lo.n] == 0) return cudaSuccess;
    if (dtype == 0) apply_fwd_nhwc<float><<<blocks_for(lo.start[lo.n]), kThreads, 0, stream>>>(p, lo);
    else apply_fwd_nhwc<__nv_bfloat16><<<blocks_for(lo.start[lo.n]), kThreads, 0, stream>>>(p, lo);
  }
  return cudaGetLastError();
}

cudaError_t launch_fpn_apply_backward(const FpnParams& p, int dtype, int layout,
                                      cudaStream_t stream) {
  if (layout == 1 && fpn_cl_ok(p, dtype, true, false) && apply_cl_ok(p, dtype) &&
      (reinterpret_cast<uintptr_t>(p.dbsf) & 15u) == 0)
    return launch_fpn_apply_backward_cl(p, dtype, stream);
  const LevelOffsets lo = offsets(p, true);
  if (lo.start[lo.n] == 0) return cudaSuccess;
  ARFE_DISPATCH(apply_bwd_gates, blocks_for(lo.start[lo.n]), p, lo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t total = (size_t)p.B * p.C * p.Hr * p.Wr;
  ARFE_DISPATCH(apply_bwd_bsf, blocks_for(total), p);
  return cudaGetLastError();
}

}  // namespace arfe
