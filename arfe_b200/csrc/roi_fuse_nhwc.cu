// Channels-last (NHWC) fast path of the AR-RFF extraction.
//
// With channels innermost every access of this gather/scatter workload is
// lane == channel coalesced, a window row of a region is one contiguous piece
// of memory, and nothing needs an atomic:
//
//   plan     : roi_prep_kernel writes, per (RoI, region), its level, window and
//              the aggregated bilinear tables (rows transposed, columns in
//              forward form) plus per-band id segments into a caller-provided
//              workspace; forward and backward share it.
//   forward  : roi_fuse_fwd_ring -- persistent CTAs, a producer warp streams
//              window rows through a shared-memory byte ring with bulk async
//              copies (TMA) + mbarriers, consumer warps (one per output column)
//              fold taps from shared memory; every window byte crosses L2 -> SM
//              once.  roi_fuse_fwd_cl (L1-cached 128-bit gathers, one CTA per
//              region) is the plan-free kernel and serves the regions the ring
//              cannot take, and NCHW output.
//   backward : PULL formulation.  roi_bin_kernel lists, per 8x8 / 8x4 tile, the
//              regions that reach it (index order) and writes one stage
//              descriptor per entry; roi_bwd_pull_tma (persistent, TMA ring)
//              streams the dout bins of each stage once per tile, warp == tile
//              row, lanes == channels, accumulators in registers -> every
//              gradient element is written exactly once, deterministically,
//              with no atomics and no zero-fill.  (The NCHW/atomic kernel is
//              bound by the L2 atomic units at ~1 fp32 element per slice-clock;
//              see DESIGN.md section 6.)
#include <stdlib.h>

#include "roi_common.cuh"
#include "tma.cuh"

namespace arfe {

// ------------------------------------------------------------------ helpers
template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int n = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int n = 8; };

template <typename T>
__device__ __forceinline__ void ldg_vec(const T* __restrict__ p, float (&f)[VecOf<T>::n]) {
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

template <typename T>
__device__ __forceinline__ void st_vec(T* __restrict__ p, const float (&f)[VecOf<T>::n]) {
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

// Packed fp32 pairs (FFMA2 / FMUL2: two IEEE fp32 operations per instruction).
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// VEC consecutive channels as packed pairs.
template <typename T>
__device__ __forceinline__ void ldg_pairs(const T* __restrict__ p, uint64_t (&f)[VecOf<T>::n / 2]) {
  if constexpr (sizeof(T) == 4) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    f[0] = v.x; f[1] = v.y;
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[i] = pack2(t.x, t.y);
    }
  }
}

// VEC consecutive channels from shared memory as packed pairs.
template <typename T>
__device__ __forceinline__ void lds_pairs(const unsigned char* __restrict__ p, uint64_t (&f)[VecOf<T>::n / 2]) {
  if constexpr (sizeof(T) == 4) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    f[0] = v.x; f[1] = v.y;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[i] = pack2(t.x, t.y);
    }
  }
}

// ------------------------------------------------------------------ forward
// One bin, VEC channels per lane: rows x NC taps, NC loads in flight per row.
// `rstep` = +1 walks the bin's rows top-down, -1 bottom-up (base / wy then
// point at the last row).
template <typename T, int NC>
__device__ __forceinline__ void cl_bin(const T* __restrict__ base, ptrdiff_t rowstride, int C,
                                       const float* __restrict__ wy, int nr, int rstep,
                                       const float* __restrict__ wx,
                                       uint64_t (&acc)[VecOf<T>::n / 2]) {
  constexpr int V2 = VecOf<T>::n / 2;
  uint64_t w[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) w[j] = pack2(wx[j], wx[j]);
#pragma unroll 2
  for (int jr = 0; jr < nr; ++jr) {
    uint64_t v[NC][V2];
#pragma unroll
    for (int j = 0; j < NC; ++j) ldg_pairs<T>(base + (size_t)j * C, v[j]);
    uint64_t t[V2];
#pragma unroll
    for (int u = 0; u < V2; ++u) t[u] = mul2(v[0][u], w[0]);
#pragma unroll
    for (int j = 1; j < NC; ++j)
#pragma unroll
      for (int u = 0; u < V2; ++u) t[u] = fma2(v[j][u], w[j], t[u]);
    const float a = *wy;
    const uint64_t ap = pack2(a, a);
#pragma unroll
    for (int u = 0; u < V2; ++u) acc[u] = fma2(t[u], ap, acc[u]);
    base += rowstride;
    wy += rstep;
  }
}

template <typename T>
__device__ __forceinline__ void cl_bin_any(const T* __restrict__ base, size_t rowstride_, int C,
                                           const float* __restrict__ wy, int nr, bool up,
                                           const float* __restrict__ wx, int nc,
                                           uint64_t (&acc)[VecOf<T>::n / 2]) {
  constexpr int V2 = VecOf<T>::n / 2;
  ptrdiff_t rowstride = (ptrdiff_t)rowstride_;
  int rstep = 1;
  if (up) {  // start at the last row and walk up
    base += (size_t)(nr - 1) * rowstride_;
    wy += nr - 1;
    rowstride = -rowstride;
    rstep = -1;
  }
  switch (nc) {
    case 1: cl_bin<T, 1>(base, rowstride, C, wy, nr, rstep, wx, acc); return;
    case 2: cl_bin<T, 2>(base, rowstride, C, wy, nr, rstep, wx, acc); return;
    case 3: cl_bin<T, 3>(base, rowstride, C, wy, nr, rstep, wx, acc); return;
    case 4: cl_bin<T, 4>(base, rowstride, C, wy, nr, rstep, wx, acc); return;
    case 5: cl_bin<T, 5>(base, rowstride, C, wy, nr, rstep, wx, acc); return;
    case 6: cl_bin<T, 6>(base, rowstride, C, wy, nr, rstep, wx, acc); return;
    default: break;
  }
  for (int jr = 0; jr < nr; ++jr) {
    const float a = *wy;
    for (int j = 0; j < nc; ++j) {
      uint64_t v[V2];
      ldg_pairs<T>(base + (size_t)j * C, v);
      const float w = a * wx[j];
      const uint64_t wp = pack2(w, w);
#pragma unroll
      for (int u = 0; u < V2; ++u) acc[u] = fma2(v[u], wp, acc[u]);
    }
    base += rowstride;
    wy += rstep;
  }
}

// Direct (L1-cached loads) channels-last forward of one region:
// warp == (bin row ph, channel chunk): the warps of neighbouring bin rows run
// side by side and walk their rows in opposite directions (even ph top-down,
// odd ph bottom-up), so the feature row two bin rows share is touched by
// both at about the same time and the second touch hits L1.
template <typename T>
__device__ void fwd_direct_cl(const RoiFuseParams& p, const CtaHeader& hd, const AxisTable& ty,
                              const AxisTable& tx, int k, int r, int nwarps) {
  constexpr int V = VecOf<T>::n;
  constexpr int V2 = V / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  const int PH = p.PH, PW = p.PW, C = p.C, BS = p.bin_stride;
  const int H = hd.H, W = hd.W;
  const float inv_count = 1.0f / hd.g.count;
  const uint64_t inv2 = pack2(inv_count, inv_count);
  const T* __restrict__ fimg =
      static_cast<const T*>(p.feats[hd.lvl]) + (size_t)hd.g.batch * H * W * C;
  const size_t rowstride = (size_t)W * C;
  const int cw = 32 * V;  // channels per warp pass
  T* __restrict__ out = static_cast<T*>(p.out);
  const int nchunk = (C + cw - 1) / cw;
  for (int item = warp; item < PH * nchunk; item += nwarps) {
    const int chunk = item / PH, ph = item - chunk * PH;
    const int c = chunk * cw + lane * V;
    if (c >= C) continue;
    const int nr = ty.cnt[ph];
    const float* __restrict__ wy = ty.w + ty.off[ph];
    const bool up = (ph & 1) != 0;
    const T* __restrict__ rowbase = fimg + (size_t)ty.first[ph] * rowstride + c;
    T* __restrict__ o = out + ((size_t)k * PH * PW + ph * PW) * BS + p.reg_off[r] + c;
    for (int pw = 0; pw < PW; ++pw) {
      const int nc = tx.cnt[pw];
      uint64_t acc[V2];
#pragma unroll
      for (int u = 0; u < V2; ++u) acc[u] = 0ull;
      if (nr > 0 && nc > 0)
        cl_bin_any<T>(rowbase + (size_t)tx.first[pw] * C, rowstride, C, wy, nr, up,
                      tx.w + tx.off[pw], nc, acc);
      float f[V];
#pragma unroll
      for (int u = 0; u < V2; ++u) unpack2(mul2(acc[u], inv2), f[2 * u], f[2 * u + 1]);
      st_vec<T>(o + (size_t)pw * BS, f);
    }
  }
}

// dynamic smem: [CtaHeader][AxisTable y][AxisTable x][outs: bins_per_pass * opitch] (NCHW out only)
template <typename T, bool kOutCL>
__device__ __forceinline__ void fwd_region_cl(const RoiFuseParams& p, int opitch, int bins_per_pass,
                                              int region, unsigned char* smem) {
  constexpr int V = VecOf<T>::n;
  CtaHeader& hd = *reinterpret_cast<CtaHeader*>(smem);
  AxisTable& ty = *reinterpret_cast<AxisTable*>(smem + 128);
  AxisTable& tx = *reinterpret_cast<AxisTable*>(smem + 128 + sizeof(AxisTable));
  float* outs = reinterpret_cast<float*>(smem + kHdrBytes);

  const int k = region / p.R, r = region % p.R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PH = p.PH, PW = p.PW, PHW = PH * PW, C = p.C;
  T* __restrict__ out = static_cast<T*>(p.out);
  // element (bin, c) of this region's output block
  auto out_index = [&](int bin, int c) -> size_t {
    return kOutCL ? ((size_t)k * PHW + bin) * p.bin_stride + p.reg_off[r] + c
                  : (((size_t)k * p.R + r) * C + c) * PHW + bin;
  };

  const bool live = setup_cta(p, k, r, hd, ty, tx);
  if (!live || hd.overflow) {
    if (live) {  // tables did not fit: reference loop order, direct taps
      const T* __restrict__ f = static_cast<const T*>(p.feats[hd.lvl]);
      const int H = hd.H, W = hd.W;
      const RoiGeom g = hd.g;
      for (int e = tid; e < C * PHW; e += kThreads) {
        const int bin = e / C, c = e - bin * C;
        const int ph = bin / PW, pw = bin % PW;
        float acc = 0.f;
        for (int iy = 0; iy < g.grid_h; ++iy) {
          AxisTap a = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H);
          if (a.lo < 0) continue;
          for (int ix = 0; ix < g.grid_w; ++ix) {
            AxisTap b = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W);
            if (b.lo < 0) continue;
            const size_t base = (size_t)g.batch * H * W;
            acc += a.wl * b.wl * to_f(f[(base + (size_t)a.lo * W + b.lo) * C + c]) +
                   a.wl * b.wh * to_f(f[(base + (size_t)a.lo * W + b.hi) * C + c]) +
                   a.wh * b.wl * to_f(f[(base + (size_t)a.hi * W + b.lo) * C + c]) +
                   a.wh * b.wh * to_f(f[(base + (size_t)a.hi * W + b.hi) * C + c]);
          }
        }
        out[out_index(bin, c)] = from_f<T>(__fdiv_rn(acc, g.count));
      }
    } else {
      for (int e = tid; e < C * PHW; e += kThreads) {
        const int bin = e / C, c = e - bin * C;
        out[out_index(bin, c)] = from_f<T>(0.f);
      }
    }
    return;
  }

  const int H = hd.H, W = hd.W;
  const float inv_count = 1.0f / hd.g.count;
  const T* __restrict__ fimg =
      static_cast<const T*>(p.feats[hd.lvl]) + (size_t)hd.g.batch * H * W * C;
  const size_t rowstride = (size_t)W * C;
  const int cw = 32 * V;  // channels per warp pass

  constexpr int V2 = V / 2;
  const float2 icp = make_float2(inv_count, inv_count);
  const uint64_t inv2 = pack2(icp.x, icp.y);
  if constexpr (kOutCL) {
    fwd_direct_cl<T>(p, hd, ty, tx, k, r, kWarps);
  } else {
  for (int b0 = 0; b0 < PHW; b0 += bins_per_pass) {
    const int b1 = min(PHW, b0 + bins_per_pass);
    for (int bin = b0 + warp; bin < b1; bin += kWarps) {
      const int ph = bin / PW, pw = bin - ph * PW;
      const int nr = ty.cnt[ph], nc = tx.cnt[pw];
      const float* __restrict__ wy = ty.w + ty.off[ph];
      const float* __restrict__ wx = tx.w + tx.off[pw];
      const T* __restrict__ base0 = fimg + ((size_t)ty.first[ph] * W + tx.first[pw]) * C + lane * V;
      for (int c0 = 0; c0 < C; c0 += cw) {
        if (c0 + lane * V >= C) continue;
        uint64_t acc[V2];
#pragma unroll
        for (int u = 0; u < V2; ++u) acc[u] = 0ull;
        if (nr > 0 && nc > 0) cl_bin_any<T>(base0 + c0, rowstride, C, wy, nr, false, wx, nc, acc);
        float* o = outs + (size_t)(bin - b0) * opitch + c0 + lane * V;
#pragma unroll
        for (int u = 0; u < V2; u += 2) {
          const ulonglong2 r2 = make_ulonglong2(mul2(acc[u], inv2), mul2(acc[u + 1], inv2));
          *reinterpret_cast<ulonglong2*>(o + 2 * u) = r2;
        }
      }
    }
    __syncthreads();
    // per channel a run of (b1 - b0) consecutive bins
    const int run = b1 - b0;
    for (int c = warp; c < C; c += kWarps)
      for (int b = lane; b < run; b += 32)
        out[out_index(b0 + b, c)] = from_f<T>(outs[(size_t)b * opitch + c]);
    __syncthreads();
  }
  }
}

// Grid = one CTA per (RoI, region); or, behind the ring kernel (p.flag_list /
// p.flag_count: the plan's fwd_list), a fixed small grid walking the regions the
// ring kernel left out.
template <typename T, bool kOutCL, int kOcc>
__global__ void __launch_bounds__(kThreads, kOcc)
roi_fuse_fwd_cl(const RoiFuseParams p, int opitch, int bins_per_pass) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (p.flag_list) {
    const int n = *p.flag_count;
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
      fwd_region_cl<T, kOutCL>(p, opitch, bins_per_pass, p.flag_list[j], smem);
      __syncthreads();
    }
  } else {
    fwd_region_cl<T, kOutCL>(p, opitch, bins_per_pass, blockIdx.x, smem);
  }
}

// ------------------------------------------------------------ region prep
// Workspace layout (bytes, each array 256-byte aligned):
//   hdr     : N * 32                         RegionHdr
//   rowtab  : N * kTabLen * 16               TapEntry per (bin block, window row)   (transposed table)
//   colbin  : N * kMaxPool * 8               ColBin per output column pw            (forward table)
//   colw    : N * kTabLen * 4                aggregated column weights, bin-major
//   seg_ids : nblk * NK * kPrepBlock * 4     region ids per (prep block, key), index order;
//   seg_cnt : nblk * NK * 4                  key = (level, image, 8-row band of the level)
constexpr int kBandH = 8;       // rows per band == rows per pull tile
constexpr int kFwdTaps = 8;     // column taps per bin the forward ring kernel takes
constexpr int kTabLen = 256;    // per region: (row bin blocks) x (window rows) <= kTabLen, column weights <= kTabLen
constexpr int kMaxBlk = 15;     // a row may be sampled by up to 2 * kMaxBlk bins (4-bit field)
constexpr int kPrepBlock = 64;  // regions per id-segment block (== threads of roi_prep_kernel)

// Entry of window row i for bin block blk; `len` = window length of the region.
__host__ __device__ inline int tab_index(int blk, int i, int len) { return blk * len + i; }

struct __align__(16) TapEntry {
  int p0;         // first bin sampling this row
  int n;          // number of bins sampling it: 0, 1 or 2
  float w0, w1;   // their aggregated weights (carry 1/count)
};

struct __align__(8) ColBin {
  int first;      // first feature column touched by output column pw
  short cnt;      // number of consecutive columns touched (0: none)
  short off;      // offset of its weights in colw
};

constexpr int kJ = 8;           // output columns (pw) per stage
constexpr int kMaxPh = 8;       // output rows (ph) per stage
constexpr int kMaxTW = 8;       // widest pull tile

// One stage of the pull kernel = one listed region of one tile: the bins
// [ph_lo, ph_lo + nph) x [pw0, pw0 + npw) of its dout block reach the tile.
struct __align__(16) StageDesc {
  int src_off;         // dout element offset of bin (ph_lo, pw0) inside the region's block
  short nph, npw;
  int region;          // which dout block (p.reg_off[region])
  int pad;
  int4 rows[kBandH];   // per tile row: {first bin row - ph_lo (< 0: row not sampled), two bins, w0, w1}
  float cw[kJ][kMaxTW];  // per output column: its weights over the tile columns
};
static_assert(sizeof(StageDesc) == 400, "descriptor is copied as one 400-byte bulk");

struct PullWs {
  RegionHdr* hdr;
  TapEntry* rowtab;
  ColBin* colbin;
  float* colw;
  int* seg_ids;
  int* seg_cnt;
  int2* tile_desc;            // per pull tile: {pool offset, stages} or {., -1}: the inline kernel serves it
  int* counters;              // backward (zeroed per call): [0] next free pool entry, [1] inline tiles,
                              // [2] next tile of the pull kernel; plan (zeroed by prep): [3] regions for
                              // the atomic fallback kernel, [4] regions for the forward fallback kernel,
                              // [5] next region of the forward ring kernel
  int* flag_list;             // regions for the atomic fallback kernel
  int* fwd_list;              // regions the forward ring kernel cannot serve
  int* wsize;                 // per region: window rows << 16 | window columns (0: not served by the ring kernel)
  int* inline_list;           // tiles for the inline kernel
  int fwd_wlen_cap;           // widest window the forward ring kernel takes (0: no forward planned)
  StageDesc* pool;            // [pool_cap] stage descriptors
  int pool_cap;
  int nblk;
  int nkeys;                  // sum over levels of B * bands(level)
  int key_base[kMaxLevels];   // first key of level l
  int nbands[kMaxLevels];     // ceil(H_l / 8)
};


__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Tiles of the two pull passes (4- and 8-pixel-wide tiles, 8 rows).
inline bool pull_heavy(int H, int W) {
  // Narrow (8 x 4) tiles at every level: smaller stages keep more of them in flight in the ring,
  // which is what the pull kernel's speed follows (8 x 8 tiles on the two large levels: 166 us,
  // narrow everywhere: 141 us on the bench workload; profiles/r2_pull_tiles.txt).
  static const long long cut = ARFE_KNOB_ENV("ARFE_PULL_HEAVY_PX", 1 << 30);
  return (long long)H * W <= cut;
}
inline int pull_tiles(int L, int B, const int* H, const int* W) {
  int n = 0;
  for (int l = 0; l < L; ++l) {
    const int tw = pull_heavy(H[l], W[l]) ? 4 : 8;
    n += ((W[l] + tw - 1) / tw) * ((H[l] + kBandH - 1) / kBandH) * B;
  }
  return n;
}
inline int pull_pool_cap(int N) { return N * 24 > 4096 ? N * 24 : 4096; }

inline size_t pull_ws_layout(int N, int L, int B, const int* H, const int* W, unsigned char* base, PullWs* ws) {
  const int nblk = (N + kPrepBlock - 1) / kPrepBlock;
  int nkeys = 0, key_base[kMaxLevels], nbands[kMaxLevels];
  for (int l = 0; l < kMaxLevels; ++l) {
    key_base[l] = nkeys;
    nbands[l] = l < L ? (H[l] + kBandH - 1) / kBandH : 0;
    nkeys += nbands[l] * B;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_hdr = take((size_t)N * sizeof(RegionHdr));
  const size_t o_row = take((size_t)N * kTabLen * sizeof(TapEntry));
  const size_t o_cbin = take((size_t)N * kMaxPool * sizeof(ColBin));
  const size_t o_colw = take((size_t)N * kTabLen * sizeof(float));
  const size_t o_ids = take((size_t)nblk * nkeys * kPrepBlock * 4);
  const size_t o_cnt = take((size_t)nblk * nkeys * 4);
  const int ntiles = pull_tiles(L, B, H, W), cap = pull_pool_cap(N);
  const size_t o_td = take((size_t)ntiles * sizeof(int2));
  const size_t o_cur = take(32);
  const size_t o_fl = take((size_t)N * 4);
  const size_t o_fwl = take((size_t)N * 4);
  const size_t o_wsz = take((size_t)N * 4);
  const size_t o_il = take((size_t)ntiles * 4);
  const size_t o_pool = take((size_t)cap * sizeof(StageDesc));
  if (ws) {
    ws->tile_desc = reinterpret_cast<int2*>(base + o_td);
    ws->counters = reinterpret_cast<int*>(base + o_cur);
    ws->flag_list = reinterpret_cast<int*>(base + o_fl);
    ws->fwd_list = reinterpret_cast<int*>(base + o_fwl);
    ws->wsize = reinterpret_cast<int*>(base + o_wsz);
    ws->fwd_wlen_cap = 0;
    ws->inline_list = reinterpret_cast<int*>(base + o_il);
    ws->pool = reinterpret_cast<StageDesc*>(base + o_pool);
    ws->pool_cap = cap;
    ws->hdr = reinterpret_cast<RegionHdr*>(base + o_hdr);
    ws->rowtab = reinterpret_cast<TapEntry*>(base + o_row);
    ws->colbin = reinterpret_cast<ColBin*>(base + o_cbin);
    ws->colw = reinterpret_cast<float*>(base + o_colw);
    ws->seg_ids = reinterpret_cast<int*>(base + o_ids);
    ws->seg_cnt = reinterpret_cast<int*>(base + o_cnt);
    ws->nblk = nblk;
    ws->nkeys = nkeys;
    for (int l = 0; l < kMaxLevels; ++l) { ws->key_base[l] = key_base[l]; ws->nbands[l] = nbands[l]; }
  }
  return off;
}

// Transpose the row table: for every window row the bins sampling it, two
// per block (block k holds bins p0+2k, p0+2k+1).  One warp; lane = window
// row.  Returns the number of blocks used (>= 1), or 0 when blocks x window
// length exceeds kTabLen: such regions take the atomic fallback kernel.
__device__ int transpose_axis(const AxisTable& t, int P, int lo, int hi, float scale,
                              TapEntry* __restrict__ tab, int lane) {
  const int n = hi - lo + 1;
  if (n > kTabLen) return 0;
  // pass 1: the largest number of bins sampling one row
  int maxcnt = 0;
  for (int i = lane; i < n; i += 32) {
    const int row = lo + i;
    int p0 = -1, cnt = 0;
    for (int q = 0; q < P; ++q)
      if (t.cnt[q] > 0 && row >= t.first[q] && row < t.first[q] + t.cnt[q]) {
        if (p0 < 0) p0 = q;
        cnt = q - p0 + 1;
      }
    maxcnt = max(maxcnt, cnt);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, d));
  int nblk = (maxcnt + 1) / 2;
  if (nblk < 1) nblk = 1;
  if (nblk > kMaxBlk || nblk * n > kTabLen) return 0;
  // pass 2: the entries
  for (int i = lane; i < n; i += 32) {
    const int row = lo + i;
    int p0 = -1, cnt = 0;
    for (int q = 0; q < P; ++q)
      if (t.cnt[q] > 0 && row >= t.first[q] && row < t.first[q] + t.cnt[q]) {
        if (p0 < 0) p0 = q;
        cnt = q - p0 + 1;
      }
    auto w_of = [&](int q) -> float {
      const int j = row - t.first[q];
      return (q < P && t.cnt[q] > 0 && j >= 0 && j < t.cnt[q]) ? t.w[t.off[q] + j] * scale : 0.f;
    };
    for (int blk = 0; blk < nblk; ++blk) {
      const int rem = cnt - 2 * blk;
      TapEntry e;
      e.p0 = (p0 < 0 ? 0 : p0) + 2 * blk;
      e.n = rem <= 0 ? 0 : (rem > 2 ? 2 : rem);
      e.w0 = e.n > 0 ? w_of(e.p0) : 0.f;
      e.w1 = e.n > 1 ? w_of(e.p0 + 1) : 0.f;
      tab[tab_index(blk, i, n)] = e;
    }
  }
  return nblk;
}

// Blocks [0, nblk): per-level ordered id segments of kPrepBlock regions each.
// Blocks [nblk, nblk + N): tap tables of one region each.  64 threads: the table
// construction is two warps wide, so small blocks keep the SMs full of them.
constexpr int kPrepThreads = kPrepBlock;
__global__ void __launch_bounds__(kPrepThreads)
roi_prep_kernel(const RoiFuseParams p, const PullWs ws) {
  const int N = p.K * p.R;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < ws.nblk) {
    // ---- per-(level, image, band) ordered id segments of kPrepBlock regions ----
    // Each region marks the keys of the 8-row bands its window can touch in a
    // per-key bitmask (order-independent), its slot is the popcount below it:
    // lists come out in region-index order without a sort.
    extern __shared__ unsigned mask[];  // [nkeys][kPrepBlock / 32]
    constexpr int kW = kPrepBlock / 32;
    for (int e = tid; e < ws.nkeys * kW; e += kPrepThreads) mask[e] = 0u;
    __syncthreads();
    const int word = tid >> 5;
    const int i = blockIdx.x * kPrepBlock + tid;  // region id = k * R + r
    int k0 = 0, k1 = -1;                           // key range of this region
    if (i < N) {
      const int k = i / p.R, r = i - k * p.R;
      RegionBox bx = region_box(p.rois + 5 * (size_t)k, r, p.facs);
      const int lvl = (p.L == 1) ? 0 : map_roi_level(bx, p.L, p.finest_scale);
      const int b = (int)bx.b;
      if (lvl >= 0 && b >= 0 && b < p.B) {
        const RoiGeom g = roi_geometry(bx, p.scale[lvl], p.PH, p.PW, p.sampling_ratio);
        // conservative row range of the sampling window (exact bounds are in hdr)
        const float end_h = g.start_h + g.bin_h * (float)p.PH;
        int r0 = (int)floorf(g.start_h) - 1, r1 = (int)ceilf(end_h) + 1;
        r0 = max(r0, 0); r1 = min(r1, p.H[lvl] - 1);
        if (r1 >= r0 && end_h == end_h) {
          k0 = ws.key_base[lvl] + b * ws.nbands[lvl] + r0 / kBandH;
          k1 = ws.key_base[lvl] + b * ws.nbands[lvl] + r1 / kBandH;
        }
      }
    }
    for (int key = k0; key <= k1; ++key) atomicOr(&mask[key * kW + word], 1u << (tid & 31));
    __syncthreads();
    for (int key = k0; key <= k1; ++key) {
      int pos = __popc(mask[key * kW + word] & ((1u << (tid & 31)) - 1u));
      for (int w = 0; w < word; ++w) pos += __popc(mask[key * kW + w]);
      ws.seg_ids[((size_t)blockIdx.x * ws.nkeys + key) * kPrepBlock + pos] = i;
    }
    for (int key = tid; key < ws.nkeys; key += kPrepThreads) {
      int c = 0;
      for (int w = 0; w < kW; ++w) c += __popc(mask[key * kW + w]);
      ws.seg_cnt[(size_t)blockIdx.x * ws.nkeys + key] = c;
    }
    return;
  }
  // ---- tap tables of region i ----
  __shared__ CtaHeader hd;
  __shared__ AxisTable ty, tx;
  __shared__ int fit;
  const int i = blockIdx.x - ws.nblk;
  const int k = i / p.R, r = i - k * p.R;
  const bool live = setup_cta(p, k, r, hd, ty, tx);
  RegionHdr h;
  h.lvl = live ? hd.lvl : -1;
  h.batch = live ? hd.g.batch : 0;
  h.ymin = h.ymax = h.xmin = h.xmax = 0;
  h.src = i;
  h.flags = 0;
  if (live && hd.overflow) h.flags = 1;
  if (live && !hd.overflow) {
    h.ymin = hd.ymin; h.ymax = hd.ymax; h.xmin = hd.xmin; h.xmax = hd.xmax;
    const int warp = tid >> 5, lane = tid & 31;
    const int ctotal = tx.off[p.PW - 1] + tx.cnt[p.PW - 1];  // off[] is a running prefix
    if (warp == 0) {
      const int nb = transpose_axis(ty, p.PH, hd.ymin, hd.ymax, 1.0f / hd.g.count,
                                    ws.rowtab + (size_t)i * kTabLen, lane);
      if (lane == 0) fit = nb;
    } else if (ctotal <= kTabLen) {
      const int t2 = tid - 32;
      if (t2 < p.PW) {
        ColBin cb;
        cb.first = tx.first[t2]; cb.cnt = (short)tx.cnt[t2]; cb.off = (short)tx.off[t2];
        ws.colbin[(size_t)i * kMaxPool + t2] = cb;
      }
      for (int j = t2; j < ctotal; j += kPrepThreads - 32) ws.colw[(size_t)i * kTabLen + j] = tx.w[j];
    }
    __syncthreads();
    if (!fit || ctotal > kTabLen) h.flags = 1;
    else h.flags = fit << 8;  // row bin blocks
  }
  if (tid == 0) {
    int wsz = 0;
    if (live) {  // can the forward ring kernel take it?
      const int wlen = h.xmax - h.xmin + 1;
      if ((h.flags & 1) || hd.overflow || wlen > ws.fwd_wlen_cap) h.flags |= 2;
      else wsz = ((h.ymax - h.ymin + 1) << 16) | wlen;
    }
    ws.hdr[i] = h;
    ws.wsize[i] = wsz;
    if (h.flags & 1) ws.flag_list[atomicAdd(ws.counters + 3, 1)] = i;
    if ((h.flags & 2) && ws.fwd_wlen_cap > 0) ws.fwd_list[atomicAdd(ws.counters + 4, 1)] = i;
  }
}

// ------------------------------------------------------- forward (TMA ring)
// Persistent CTAs (one or two per SM) take (RoI, region)s from a work counter in
// the plan; channels-last in and out; the region tables come from the plan
// written by roi_prep_kernel (the same plan the backward uses).
// In NHWC a row of a region's sampling window (wlen pixels x C channels) is ONE
// contiguous piece of memory, so the PRODUCER warp streams the windows row by
// row into a shared-memory byte ring with bulk async copies (TMA) signalling an
// mbarrier per row, and the small tables of the next regions into one of
// kFwdTabs table buffers; it runs ahead across regions, so rows stay in flight
// without holding registers and every window byte crosses L2 -> SM exactly once
// (the L1-cached kernel above re-fetches ~1.5x).  CONSUMER warp == (output
// column pw, channel chunk): per row it folds its (<= 8) column taps into
// t = sum_j wx[j] * f[row][x0 + j] from shared memory and adds wy[ph][row] * t
// to the bin rows ph that sample the row (two per table block); its PH
// accumulators stay in registers and are written once per region.
// Regions the ring cannot serve (flag bit 1 in the plan: window row wider than
// half the ring, a column with more than 8 taps, tables that did not fit) are
// left to the L1-cached kernel, launched over the plan's fwd_list.
constexpr int kFwdSlots = 16;
constexpr int kFwdTabs = 3;
constexpr int kFwdCopy = 65536;  // bytes per bulk copy (a typical 22 KB window row is one copy)

struct __align__(128) FwdTab {
  RegionHdr hdr;
  int pad[8];
  TapEntry rowtab[kTabLen];
  ColBin colbin[kMaxPool];
  float colw[kTabLen];
};

struct FwdPipe {
  uint64_t* full;
  uint64_t* empty;
  const uint32_t* stage_off;
  const unsigned char* ring;
};

// acc[P0] += w0 * t, acc[P0 + 1] += w1 * t with P0 a run-time value: the
// accumulators live in registers, so the bin row is resolved by a jump table.
// (Measured alternatives for PH = 7, forward call 175 us: updating all 7 rows with a
// selected weight, 0 for the rows not sampled -- 170 us, but 0 * NaN/Inf would leak
// into bins that do not sample the pixel; one predicated FFMA2 per row -- 181 us.)
template <int PH, int V2>
__device__ __forceinline__ void add_rows(uint64_t (&acc)[PH][V2], int p0, const uint64_t (&t)[V2],
                                         float w0, float w1) {
  const uint64_t w0p = pack2(w0, w0), w1p = pack2(w1, w1);
#define ARFE_ROW_CASE(P)                                                         \
  case P:                                                                        \
    if constexpr (P < PH) {                                                      \
      _Pragma("unroll") for (int u = 0; u < V2; ++u) {                           \
        acc[P][u] = fma2(t[u], w0p, acc[P][u]);                                  \
        if constexpr (P + 1 < PH) acc[P + 1][u] = fma2(t[u], w1p, acc[P + 1][u]); \
      }                                                                          \
    }                                                                            \
    break;
  switch (p0) {
    ARFE_ROW_CASE(0) ARFE_ROW_CASE(1) ARFE_ROW_CASE(2) ARFE_ROW_CASE(3) ARFE_ROW_CASE(4)
    ARFE_ROW_CASE(5) ARFE_ROW_CASE(6) ARFE_ROW_CASE(7) ARFE_ROW_CASE(8) ARFE_ROW_CASE(9)
    ARFE_ROW_CASE(10) ARFE_ROW_CASE(11) ARFE_ROW_CASE(12) ARFE_ROW_CASE(13)
    default: break;
  }
#undef ARFE_ROW_CASE
}

// One consumer warp, one region: NC column taps per row (compile-time), NCH
// channel chunks of 32 * V channels per lane set, rows streamed through the
// ring from stage `stage0` on.
template <typename T, int PH, int NC, int NCH>
__device__ __forceinline__ void fwd_consume_rows(const FwdPipe& pp, int stage0, int nrows, int nblk,
                                                 const TapEntry* __restrict__ rowtab,
                                                 const float* __restrict__ wxp, uint32_t tap0, int C,
                                                 T* __restrict__ o, size_t ostep, int nc_dyn = 0) {
  constexpr int V = VecOf<T>::n;
  constexpr int V2 = V / 2;
  constexpr uint32_t kChunkBytes = 32 * V * sizeof(T);
  const int lane = threadIdx.x & 31;
  const uint32_t tap_step = (uint32_t)C * sizeof(T);
  float wx[NC > 0 ? NC : 1];
#pragma unroll
  for (int j = 0; j < NC; ++j) wx[j] = wxp[j];
  if (nblk == 1) {
    // Bins at least one pixel high: a window row is sampled by at most the two adjacent bin
    // rows p0, p0 + 1 and p0 never decreases down the window, so two live accumulators per
    // chunk suffice.  A bin row is written out as soon as the window has moved past it
    // (same additions in the same order as the general path below: identical bits) -- no
    // run-time index into register accumulators, hence no jump table.
    uint64_t lo[NCH][V2], hi[NCH][V2];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int u = 0; u < V2; ++u) lo[ch][u] = hi[ch][u] = 0ull;
    int cur = 0;  // bin row held in lo; hi holds cur + 1
    auto flush = [&]() {
      if (o) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float f[V];
#pragma unroll
          for (int u = 0; u < V2; ++u) unpack2(lo[ch][u], f[2 * u], f[2 * u + 1]);
          st_vec<T>(o + (size_t)cur * ostep + ch * (32 * V), f);
        }
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int u = 0; u < V2; ++u) { lo[ch][u] = hi[ch][u]; hi[ch][u] = 0ull; }
      ++cur;
    };
    for (int rr = 0; rr < nrows; ++rr) {
      const int stage = stage0 + rr;
      const int slot = stage % kFwdSlots;
      const int4 rec = *reinterpret_cast<const int4*>(rowtab + rr);
      mbar_wait(pp.full + slot, (stage / kFwdSlots) & 1);
      if (NC > 0 && rec.y > 0) {
        while (cur < rec.x) flush();  // warp-uniform
        const uint64_t w0p = pack2(__int_as_float(rec.z), __int_as_float(rec.z));
        const uint64_t w1p = pack2(__int_as_float(rec.w), __int_as_float(rec.w));
        const unsigned char* __restrict__ s0 = pp.ring + pp.stage_off[slot] + tap0;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const unsigned char* __restrict__ sc = s0 + ch * kChunkBytes;
          uint64_t v[NC > 0 ? NC : 1][V2];
#pragma unroll
          for (int j = 0; j < NC; ++j) lds_pairs<T>(sc + j * tap_step, v[j]);
          uint64_t t[V2];
#pragma unroll
          for (int u = 0; u < V2; ++u) t[u] = mul2(v[0][u], pack2(wx[0], wx[0]));
#pragma unroll
          for (int j = 1; j < NC; ++j)
#pragma unroll
            for (int u = 0; u < V2; ++u) t[u] = fma2(v[j][u], pack2(wx[j], wx[j]), t[u]);
          for (int j = NC; j < nc_dyn; ++j) {
            uint64_t vv[V2];
            lds_pairs<T>(sc + j * tap_step, vv);
            const float w = wxp[j];
#pragma unroll
            for (int u = 0; u < V2; ++u) t[u] = fma2(vv[u], pack2(w, w), t[u]);
          }
#pragma unroll
          for (int u = 0; u < V2; ++u) {
            lo[ch][u] = fma2(t[u], w0p, lo[ch][u]);
            hi[ch][u] = fma2(t[u], w1p, hi[ch][u]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pp.empty + slot);
    }
    while (cur < PH) flush();
    return;
  }
  uint64_t acc[NCH][PH][V2];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int ph = 0; ph < PH; ++ph)
#pragma unroll
      for (int u = 0; u < V2; ++u) acc[ch][ph][u] = 0ull;
  for (int rr = 0; rr < nrows; ++rr) {
    const int stage = stage0 + rr;
    const int slot = stage % kFwdSlots;
    const int4 rec = *reinterpret_cast<const int4*>(rowtab + rr);
    mbar_wait(pp.full + slot, (stage / kFwdSlots) & 1);
    if (NC > 0 && rec.y > 0) {
      const unsigned char* __restrict__ s0 = pp.ring + pp.stage_off[slot] + tap0;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const unsigned char* __restrict__ sc = s0 + ch * kChunkBytes;
        uint64_t v[NC > 0 ? NC : 1][V2];
#pragma unroll
        for (int j = 0; j < NC; ++j) lds_pairs<T>(sc + j * tap_step, v[j]);
        uint64_t t[V2];
#pragma unroll
        for (int u = 0; u < V2; ++u) t[u] = mul2(v[0][u], pack2(wx[0], wx[0]));
#pragma unroll
        for (int j = 1; j < NC; ++j)
#pragma unroll
          for (int u = 0; u < V2; ++u) t[u] = fma2(v[j][u], pack2(wx[j], wx[j]), t[u]);
        for (int j = NC; j < nc_dyn; ++j) {  // more than kFwdTaps taps: weights from shared memory
          uint64_t vv[V2];
          lds_pairs<T>(sc + j * tap_step, vv);
          const float w = wxp[j];
#pragma unroll
          for (int u = 0; u < V2; ++u) t[u] = fma2(vv[u], pack2(w, w), t[u]);
        }
        add_rows<PH, V2>(acc[ch], rec.x, t, __int_as_float(rec.z), __int_as_float(rec.w));
        for (int rb = 1; rb < nblk; ++rb) {  // sub-pixel bins: more than two bin rows sample this row
          const int4 r2 = *reinterpret_cast<const int4*>(rowtab + tab_index(rb, rr, nrows));
          if (r2.y > 0) add_rows<PH, V2>(acc[ch], r2.x, t, __int_as_float(r2.z), __int_as_float(r2.w));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(pp.empty + slot);
  }
  if (o) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int ph = 0; ph < PH; ++ph) {
        float f[V];
#pragma unroll
        for (int u = 0; u < V2; ++u) unpack2(acc[ch][ph][u], f[2 * u], f[2 * u + 1]);
        st_vec<T>(o + (size_t)ph * ostep + ch * (32 * V), f);
      }
  }
}

// dynamic smem: [FwdTab x kFwdTabs][barriers, stage offsets: 512 bytes][ring]
// NCH = channel chunks per consumer warp: 2 halves the barrier operations per
// byte (the per-row wait / arrive round trip, not bandwidth, bounds the pipeline).
// NT = threads per CTA (256: up to 7 consumer warps, two CTAs per SM at 128
// registers; 480: up to 14, two CTAs per SM only for 7x7 fp32 single chunks).
template <typename T, int PH, int NCH, int NT>
__global__ void __launch_bounds__(NT, ((NT == 256 || PH * VecOf<T>::n * NCH <= 32) ? 2 : 1))
roi_fuse_fwd_ring(const RoiFuseParams p, const PullWs ws, int ncons, int ring_bytes) {
  constexpr int V = VecOf<T>::n;
  extern __shared__ __align__(16) unsigned char smem[];
  FwdTab* tabs = reinterpret_cast<FwdTab*>(smem);
  unsigned char* ctl = smem + kFwdTabs * sizeof(FwdTab);
  uint64_t* full = reinterpret_cast<uint64_t*>(ctl);
  uint64_t* empty = full + kFwdSlots;
  uint64_t* tab_full = empty + kFwdSlots;
  uint64_t* tab_empty = tab_full + kFwdTabs;
  uint32_t* stage_off = reinterpret_cast<uint32_t*>(tab_empty + kFwdTabs);
  unsigned char* ring = ctl + 512;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PW = p.PW, PHW = PH * PW, C = p.C, BS = p.bin_stride;
  const int N = p.K * p.R;
  T* __restrict__ out = static_cast<T*>(p.out);

  if (tid == 0) {
    for (int i = 0; i < kFwdSlots; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, ncons); }
    for (int i = 0; i < kFwdTabs; ++i) { mbar_init(tab_full + i, 1); mbar_init(tab_empty + i, ncons); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == ncons) {
    // ------------------------------------------------------------ producer
    uint32_t head = 0, tail = 0;  // live bytes of the ring: [tail, head) modulo wrap
    uint32_t my_off = 0;          // lane j: ring offset of the stage in slot j
    int issued = 0, released = 0; // stages (rows) issued / known to be consumed
    // Regions are handed out dynamically from the plan's counter.  The claim of
    // region i + 2 and the table copies of region i + 1 (header, row table,
    // column table: fixed 5.4 KB straight from the plan) are issued before the
    // rows of region i, so neither the atomic nor the header fetch sits on the
    // row stream's critical path.
    auto claim = [&]() -> int { return lane == 0 ? atomicAdd(ws.counters + 5, 1) : 0; };
    auto issue_tables = [&](int reg, int j) {  // tables of the j-th region of this CTA
      const int buf = j % kFwdTabs;
      if (j >= kFwdTabs) mbar_wait(tab_empty + buf, ((j / kFwdTabs) - 1) & 1);
      FwdTab& tb = tabs[buf];
      if (lane == 0) {
        if (reg >= N) {  // no more work: tell the consumers (and ourselves)
          tb.hdr.lvl = -9;
          mbar_arrive(tab_full + buf);
        } else {
          mbar_arrive_expect_tx(tab_full + buf, sizeof(RegionHdr) + sizeof(tb.rowtab) + sizeof(tb.colbin) + sizeof(tb.colw));
          bulk_g2s(&tb.hdr, ws.hdr + reg, sizeof(RegionHdr), tab_full + buf);
          bulk_g2s(tb.rowtab, ws.rowtab + (size_t)reg * kTabLen, sizeof(tb.rowtab), tab_full + buf);
          bulk_g2s(tb.colbin, ws.colbin + (size_t)reg * kMaxPool, sizeof(tb.colbin), tab_full + buf);
          bulk_g2s(tb.colw, ws.colw + (size_t)reg * kTabLen, sizeof(tb.colw), tab_full + buf);
        }
      }
    };
    int c_next = claim();
    issue_tables(__shfl_sync(0xffffffffu, c_next, 0), 0);
    c_next = claim();
    for (int i = 0;; ++i) {
      const int reg_next = __shfl_sync(0xffffffffu, c_next, 0);
      c_next = claim();
      issue_tables(reg_next, i + 1);
      const int buf = i % kFwdTabs;
      mbar_wait(tab_full + buf, (i / kFwdTabs) & 1);
      const RegionHdr h = tabs[buf].hdr;
      if (h.lvl == -9) break;
      const bool ringed = h.lvl >= 0 && (h.flags & 3) == 0;
      const int nrows = h.ymax - h.ymin + 1, wlen = h.xmax - h.xmin + 1;
      if (!ringed) continue;
      const uint32_t bytes = (uint32_t)wlen * C * sizeof(T);
      const unsigned char* __restrict__ src = reinterpret_cast<const unsigned char*>(
          static_cast<const T*>(p.feats[h.lvl]) +
          (((size_t)h.batch * p.H[h.lvl] + h.ymin) * p.W[h.lvl] + h.xmin) * C);
      const size_t src_step = (size_t)p.W[h.lvl] * C * sizeof(T);
      // Rows of one region are equally long, so the producer hands out ring space
      // to as many rows as fit at once: lane j issues row rr + j (its barrier arm and
      // its bulk copies).  When the producer has fallen behind, several rows are
      // free and one pass of this loop catches up -- its ~60-instruction dependent
      // path per pass, not bandwidth, was what the consumers waited for.
      auto retire = [&]() {  // `released` was consumed: its space is free
        ++released;
        const uint32_t nxt = __shfl_sync(0xffffffffu, my_off, released % kFwdSlots);
        tail = released < issued ? nxt : head;
      };
      for (int rr = 0; rr < nrows;) {
        while (released < issued && mbar_test(empty + (released % kFwdSlots), (released / kFwdSlots) & 1)) retire();
        int m = 0;
        while (true) {
          if (released == issued) head = tail = 0;                                   // ring empty
          uint32_t room = head >= tail ? (uint32_t)ring_bytes - head : tail - head - 1u;
          if (head >= tail && room < bytes && bytes < tail) { head = 0; room = tail - 1u; }  // wrap
          m = min(min((int)(room / bytes), nrows - rr), kFwdSlots - (issued - released));
          if (m > 0) break;
          mbar_wait(empty + (released % kFwdSlots), (released / kFwdSlots) & 1);     // wait for the oldest row
          retire();
        }
        {  // lane s < kFwdSlots keeps the ring offset of the stage in slot s
          const int j = (lane - issued % kFwdSlots + kFwdSlots) % kFwdSlots;
          if (lane < kFwdSlots && j < m) my_off = head + (uint32_t)j * bytes;
        }
        if (lane < m) {
          const int slot = (issued + lane) % kFwdSlots;
          const uint32_t off = head + (uint32_t)lane * bytes;
          const unsigned char* __restrict__ rsrc = src + (size_t)(rr + lane) * src_step;
          stage_off[slot] = off;
          if (ARFE_SKIP(p, 2)) {  // profiling aid: no copies
            mbar_arrive(full + slot);
          } else {
            mbar_arrive_expect_tx(full + slot, bytes);
            for (uint32_t o = 0; o < bytes; o += kFwdCopy)
              bulk_g2s(ring + off + o, rsrc + o, min((uint32_t)kFwdCopy, bytes - o), full + slot);
          }
        }
        __syncwarp();
        head += (uint32_t)m * bytes;
        issued += m;
        rr += m;
      }
    }
    return;
  }
  if (warp > ncons) return;

  // ---------------------------------------------------------------- consumers
  const int cwid = 32 * V * NCH;
  const int pw = warp % PW, chunk = warp / PW;
  const int c = chunk * cwid + lane * V;
  const bool act = c < C;  // NCH == 2 requires C % (64 * V) == 0 (launcher)
  int stage = 0;
  for (int i = 0;; ++i) {
    const int buf = i % kFwdTabs;
    mbar_wait(tab_full + buf, (i / kFwdTabs) & 1);
    const FwdTab& tb = tabs[buf];
    const int lvl = tb.hdr.lvl, flags = tb.hdr.flags;
    if (lvl == -9) break;  // no more work
    const int reg = tb.hdr.src;
    const int k = reg / p.R, r = reg - k * p.R;
    if (lvl >= 0 && (flags & 3) == 0) {
      const int nrows = tb.hdr.ymax - tb.hdr.ymin + 1;
      const ColBin cb = tb.colbin[pw];
      const int nc = cb.cnt;
      const float* __restrict__ wxp = tb.colw + cb.off;
      const uint32_t tap0 = (uint32_t)((nc > 0 ? cb.first - tb.hdr.xmin : 0) * C + (act ? c : 0)) * sizeof(T);
      T* __restrict__ o = act ? out + ((size_t)k * PHW + pw) * BS + p.reg_off[r] + c : nullptr;
      const int nblk = (flags >> 8) & 15;
      FwdPipe pipe{full, empty, stage_off, ring};
#define ARFE_CONSUME(NCC) \
  fwd_consume_rows<T, PH, NCC, NCH>(pipe, stage, nrows, nblk, tb.rowtab, wxp, tap0, C, o, (size_t)PW * BS)
      switch (ARFE_SKIP(p, 1) ? 0 : nc) {  // profiling aid: 0 taps == no math
        case 0: ARFE_CONSUME(0); break;
        case 1: ARFE_CONSUME(1); break;
        case 2: ARFE_CONSUME(2); break;
        case 3: ARFE_CONSUME(3); break;
        case 4: ARFE_CONSUME(4); break;
        case 5: ARFE_CONSUME(5); break;
        case 6: ARFE_CONSUME(6); break;
        case 7: ARFE_CONSUME(7); break;
        case 8: ARFE_CONSUME(8); break;
        default:
          fwd_consume_rows<T, PH, 8, NCH>(pipe, stage, nrows, nblk, tb.rowtab, wxp, tap0, C, o, (size_t)PW * BS, nc);
          break;
      }
#undef ARFE_CONSUME
      stage += nrows;
    } else if (lvl < 0) {  // nothing is pooled: zeros
      T* __restrict__ o = out + (size_t)k * PHW * BS + p.reg_off[r];
      for (int e = tid; e < C * PHW; e += ncons * 32) {
        const int bin = e / C, cc = e - bin * C;
        o[(size_t)bin * BS + cc] = from_f<T>(0.f);
      }
    }  // else: served by the L1-cached kernel over ws.fwd_list
    __syncwarp();
    if (lane == 0) mbar_arrive(tab_empty + buf);
  }
}

// ----------------------------------------------------------- pull backward
constexpr int kTileH = kBandH;  // tile rows, warp == tile row
constexpr int kListCap = 512;   // list entries per pass
constexpr int kChunkR = 32;     // list entries expanded per round
constexpr int kMaxPrepBlocks = 2047;  // K * regions <= 131 008 for the pull path
static_assert(kTileH * kChunkR == kThreads && kChunkR * kJ == kThreads, "expand: one item per thread");

struct TileMap {
  int tstart[kMaxLevels + 1];  // first tile of each scheduled slot
  int level[kMaxLevels];       // level of slot j (heaviest first)
  int tiles_x[kMaxLevels], tiles_y[kMaxLevels];
  int nslots;
  int groups;                  // channel groups per tile (C / (32 * V * NV))
  int tile_base;               // first tile_desc slot of this pass
};

// Compact copy of a listed (region, row bin block, pw block) in shared memory.
struct __align__(16) ListEntry {
  int id;            // region id | row block << 24
  int src;           // k * R + r
  short ymin, ymax;  // window rows
  short pw0, npw;    // output columns [pw0, pw0 + npw) reach this tile's columns
};

struct Tile {
  int l, b, y0, x0, y1, x1, key, H, W;
};

__device__ __forceinline__ Tile decode_tile(const RoiFuseParams& p, const PullWs& ws,
                                            const TileMap& tm, int t, int TW) {
  int j = 0;
  while (j + 1 < tm.nslots && t >= tm.tstart[j + 1]) ++j;
  Tile T;
  T.l = tm.level[j];
  t -= tm.tstart[j];
  const int per_img = tm.tiles_x[j] * tm.tiles_y[j];
  T.b = t / per_img;
  t -= T.b * per_img;
  const int tyi = t / tm.tiles_x[j], txi = t - tyi * tm.tiles_x[j];
  T.y0 = tyi * kTileH;
  T.x0 = txi * TW;
  T.H = p.H[T.l];
  T.W = p.W[T.l];
  T.y1 = min(T.y0 + kTileH, T.H) - 1;
  T.x1 = min(T.x0 + TW, T.W) - 1;
  T.key = ws.key_base[T.l] + T.b * ws.nbands[T.l] + tyi;  // this tile's (level, image, band)
  return T;
}

struct ListSmem {
  ListEntry list[kListCap];
  int pre[kMaxPrepBlocks + 1];
  int warp_tot[kThreads / 32];
  int list_n;
};

// candidates = the ids of this tile's (level, image, band) key over all prep
// blocks, block-major == region-index order; pre[] = exclusive prefix of the
// per-block counts so that one 256-thread batch spans blocks.  CTA-wide.
__device__ int init_candidates(ListSmem& sm, const PullWs& ws, int key) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i <= ws.nblk; i += (int)blockDim.x)
    sm.pre[i] = i < ws.nblk ? ws.seg_cnt[(size_t)i * ws.nkeys + key] : 0;
  __syncthreads();
  if (warp == 0) {  // exclusive scan of pre[0..nblk] by one warp
    int carry = 0;
    for (int i0 = 0; i0 <= ws.nblk; i0 += 32) {
      const int i = i0 + lane;
      const int v = i <= ws.nblk ? sm.pre[i] : 0;
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t2 = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t2;
      }
      if (i <= ws.nblk) sm.pre[i] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  return sm.pre[ws.nblk];
}

// Ordered compaction of the regions intersecting the tile into sm.list, from
// candidate `pos` on, until the candidates or the list capacity run out.
// Returns the new cursor.  CTA-wide (uniform control flow).
// jcap: output columns per list entry (narrow tiles: 4, so that a stage stays small enough for
// two producer rings at three CTAs per SM; wide tiles: kJ).
__device__ int build_list(ListSmem& sm, const RoiFuseParams& p, const PullWs& ws, const Tile& T,
                          int pos, int total_cand, int jcap) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PW = p.PW;
  __syncthreads();
  if (tid == 0) sm.list_n = 0;
  __syncthreads();
  while (pos < total_cand) {
    const int q = pos + tid;
    int mine = 0, nbr = 0, ncb = 0, id = 0, plo = 0, phi = -1;
    RegionHdr h;
    if (q < total_cand) {
      int lo = 0, hi = ws.nblk - 1;  // last block with pre[blk] <= q
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sm.pre[mid] <= q) lo = mid; else hi = mid - 1;
      }
      id = ws.seg_ids[((size_t)lo * ws.nkeys + T.key) * kPrepBlock + (q - sm.pre[lo])];
      h = ws.hdr[id];
      const bool hit = h.lvl == T.l && h.batch == T.b && (h.flags & 1) == 0 && h.ymax >= T.y0 &&
                       h.ymin <= T.y1 && h.xmax >= T.x0 && h.xmin <= T.x1;
      if (hit) {
        // output columns whose samples reach [x0, x1]
        const ColBin* __restrict__ cbp = ws.colbin + (size_t)id * kMaxPool;
        plo = PW;
        for (int pw = 0; pw < PW; ++pw) {
          const ColBin c = cbp[pw];
          if (c.cnt > 0 && c.first <= T.x1 && c.first + c.cnt - 1 >= T.x0) {
            plo = min(plo, pw);
            phi = pw;
          }
        }
        if (phi >= 0) {
          nbr = (h.flags >> 8) & 15;
          ncb = (phi - plo + jcap) / jcap;
          mine = nbr * ncb;
        }
      }
    }
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    __syncthreads();
    int base = sm.list_n;
    for (int w = 0; w < warp; ++w) base += sm.warp_tot[w];
    // a candidate is accepted while its entries still fit the list: acceptance
    // is monotone in candidate order, so the accepted ones form a prefix
    const int end = base + incl;
    const bool fits = end <= kListCap;
    const int accepted = __syncthreads_count(q < total_cand && fits);  // also: everyone has read list_n
    if (fits && mine) {
      ListEntry e;
      e.src = h.src;
      e.ymin = (short)h.ymin; e.ymax = (short)h.ymax;
      int o = end - mine;
      for (int rb = 0; rb < nbr; ++rb)
        for (int cb = 0; cb < ncb; ++cb) {
          e.id = id | (rb << 24);
          e.pw0 = (short)(plo + cb * jcap);
          e.npw = (short)min(jcap, phi - (plo + cb * jcap) + 1);
          sm.list[o++] = e;
        }
      atomicMax(&sm.list_n, end);
    }
    const int batch = min((int)blockDim.x, total_cand - pos);
    pos += accepted;
    __syncthreads();
    if (accepted < batch) break;  // list full
  }
  __syncthreads();
  return pos;
}

// Row record of list entry e for feature row yy: {first bin row p0 (< 0: row
// not sampled), two bins, w0, w1}.
__device__ __forceinline__ int4 expand_row(const PullWs& ws, const ListEntry& e, int yy) {
  int4 d = make_int4(-1, 0, 0, 0);
  if (yy >= e.ymin && yy <= e.ymax) {
    const int id = e.id & 0xffffff, rb = e.id >> 24;
    const TapEntry re = ws.rowtab[(size_t)id * kTabLen + tab_index(rb, yy - e.ymin, e.ymax - e.ymin + 1)];
    if (re.n > 0) {
      d.x = re.p0;
      d.y = re.n > 1 ? 1 : 0;
      d.z = __float_as_int(re.w0);
      d.w = __float_as_int(re.w1);
    }
  }
  return d;
}

// Column weights of output column e.pw0 + jj over the tile columns x0 .. x0 + TW - 1.
template <int TW>
__device__ __forceinline__ void expand_col(const PullWs& ws, const ListEntry& e, int jj, int x0,
                                           float (&w)[TW]) {
  const int id = e.id & 0xffffff;
  const ColBin c = ws.colbin[(size_t)id * kMaxPool + e.pw0 + jj];
  const float* __restrict__ cwp = ws.colw + (size_t)id * kTabLen + c.off;
#pragma unroll
  for (int x = 0; x < TW; ++x) {
    const int i = x0 + x - c.first;
    w[x] = (i >= 0 && i < c.cnt) ? __ldg(cwp + i) : 0.f;
  }
}

// The 8 tile rows of entry q live in 8 consecutive lanes: the bin-row range
// [lo, hi] they reference.
__device__ __forceinline__ void row_range(const int4 r, int& lo, int& hi) {
  lo = r.x >= 0 ? r.x : (1 << 20);
  hi = r.x >= 0 ? r.x + r.y : -1;
#pragma unroll
  for (int d = 1; d < kTileH; d <<= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
}

// Stage descriptor of list entry e for the tile, row part: thread == (entry,
// row), [lo, hi] from row_range(); (entry, pw) threads add the column weights
// separately.  Returns false when the entry references more
// than kMaxPh bin rows (the tile then goes to the inline kernel).
__device__ __forceinline__ bool write_stage_rows(const RoiFuseParams& p, StageDesc* sd,
                                                 const ListEntry& e, int4 r, int row, int lo, int hi) {
  const int nph = hi >= lo ? hi - lo + 1 : 0;
  if (r.x >= 0) r.x -= lo;
  sd->rows[row] = r;
  if (row == 0) {
    const int k = e.src / p.R, rr = e.src - k * p.R;
    sd->src_off = nph > 0 ? (k * p.PH * p.PW + lo * p.PW + e.pw0) * p.bin_stride : 0;
    sd->nph = (short)nph;
    sd->npw = e.npw;
    sd->region = rr;
  }
  return nph <= kMaxPh;
}

// ---- binning kernel: per tile, the ordered list of regions that reach it and
// one stage descriptor per list entry, written to a pool in the workspace (one
// atomic allocation per tile; the pool position varies from run to run, the
// contents and their order do not).  Tiny register footprint -> the
// latency-bound table walking runs at full occupancy, off the streaming
// kernel's critical path.  Tiles the pool cannot take (it is exhausted, or an
// entry spans more than kMaxPh bin rows) are appended to the inline list:
// roi_bwd_pull_inline builds their lists itself.
template <int TW>
__device__ void bin_tile(ListSmem& sm, const RoiFuseParams& p, const PullWs& ws,
                         const TileMap& tm, int t) {
  __shared__ int s_off;
  const int tid = threadIdx.x;
  const Tile T = decode_tile(p, ws, tm, t, TW);
  const int total_cand = init_candidates(sm, ws, T.key);
  // How many list entries in all?  Usually one pass of the list builder holds
  // them; long lists (upper levels, thousands of RoIs) are counted pass by pass
  // first, so that the tile still gets ONE contiguous run of the pool.
  constexpr int jcap = TW == 4 ? 4 : kJ;
  int pos = build_list(sm, p, ws, T, 0, total_cand, jcap);
  int n = sm.list_n;
  const bool single = pos >= total_cand;
  while (pos < total_cand) {
    pos = build_list(sm, p, ws, T, pos, total_cand, jcap);
    n += sm.list_n;
  }
  if (tid == 0) {
    int off = 0;
    if (n > 0) {
      off = atomicAdd(ws.counters, n);
      if (off + n > ws.pool_cap) off = -1;  // pool exhausted: inline kernel
    }
    s_off = off;
  }
  __syncthreads();
  const int off = s_off;
  bool ok = off >= 0;
  if (ok && n > 0) {
    const int chunk = (int)blockDim.x / kTileH;
    int done = 0;  // entries written so far
    pos = 0;
    do {
      if (!single) pos = build_list(sm, p, ws, T, pos, total_cand, jcap);  // else: the list is still in place
      const int m = sm.list_n;
      for (int c0 = 0; c0 < m; c0 += chunk) {
        const int nc = min(chunk, m - c0);
        {  // thread == (entry, tile row)
          const int q = tid / kTileH, row = tid - q * kTileH;
          const int qq = q < nc ? q : 0;  // idle lanes shadow entry 0 (no stores)
          const ListEntry e = sm.list[c0 + qq];
          const int4 r = expand_row(ws, e, T.y0 + row);
          int lo, hi;
          row_range(r, lo, hi);  // all lanes: the shuffles stay convergent
          if (q < nc) ok = write_stage_rows(p, ws.pool + off + done + c0 + q, e, r, row, lo, hi) && ok;
        }
        {  // thread == (entry, pw)
          const int q = tid / kJ, jj = tid - q * kJ;
          if (q < nc) {
            const ListEntry e = sm.list[c0 + q];
            if (jj < e.npw) {
              float w[TW];
              expand_col<TW>(ws, e, jj, T.x0, w);
              float* o = ws.pool[off + done + c0 + q].cw[jj];
#pragma unroll
              for (int x = 0; x < TW; x += 4)
                *reinterpret_cast<float4*>(o + x) = make_float4(w[x], w[x + 1], w[x + 2], w[x + 3]);
            }
          }
        }
      }
      done += m;
    } while (!single && pos < total_cand);
  }
  ok = __syncthreads_and(ok);
  if (tid == 0) {
    ws.tile_desc[tm.tile_base + t] = make_int2(off, ok ? n : -1);
    if (!ok) ws.inline_list[atomicAdd(ws.counters + 1, 1)] = tm.tile_base + t;
  }
}

constexpr int kBinThreads = 128;
static_assert(kJ == kTileH, "bin_tile: (entry, row) and (entry, pw) share the thread mapping");
__global__ void __launch_bounds__(kBinThreads)
roi_bin_kernel(const RoiFuseParams p, const PullWs ws, const TileMap tm4, const TileMap tm8, int nt4) {
  __shared__ ListSmem sm;
  if ((int)blockIdx.x < nt4) bin_tile<4>(sm, p, ws, tm4, blockIdx.x);
  else bin_tile<8>(sm, p, ws, tm8, blockIdx.x - nt4);
}

// ------------------------------------------------------------ pull kernel
// CTA == (tile of 8 rows x TW columns, group of 32 * V channels); 8 consumer
// warps (warp == tile row, lanes == channels) + 1 producer warp.
//
// The bilinear sum is separable: a gradient pixel (y, x) of a region receives
//     sum_ph ry[ph][y] * sum_pw cx[pw][x] * dout[ph][pw]
// with ry / cx the aggregated row / column weights.  For every listed region
// (stage) the PRODUCER copies the bins [ph_lo, +nph) x [pw0, +npw) of the
// region's dout block that reach the tile -- each bin's channel group is one
// contiguous piece -- plus the 400-byte stage descriptor into a shared-memory
// byte ring with bulk async copies (TMA) signalling the stage's mbarrier; it
// runs up to kNSlot stages / kRing bytes ahead, so the loads of many regions
// are in flight without holding registers.  Every bin is fetched once per tile
// and shared by the 8 rows.  A CONSUMER warp combines, per output column, the
// (<= 2) bins sampling its row with the row weights and adds the result to its
// TW pixels with that column's weights (zero where it does not reach): FFMA2
// arithmetic on shared-memory operands, accumulators in registers, every
// gradient element written once -- no atomics, no zero-fill, deterministic.
constexpr int kNSlot = 16;      // stages (groups of list entries) in flight
constexpr int kDescSlots = 32;  // list-entry descriptors in flight
constexpr int kMaxGroup = 8;    // list entries per stage
constexpr int kDescBytes = (int)sizeof(StageDesc);
constexpr int kTileQ = 4;      // tiles announced ahead of the consumers
constexpr int kPullCtl = 1024; // barriers, stage offsets, tile queue

// The kernel is PERSISTENT: CTAs (three per SM) claim (tile, channel group)s from
// a work counter, heavy tiles first; the producer announces each tile to the
// consumers through a small queue and keeps issuing the next tile's stages
// while the consumers finish the current one, so a tile's start-up latency
// (claim -> tile descriptor -> stage headers -> first bins) is hidden behind
// the previous tile instead of being paid 14 times per SM slot.
struct PullCtl {
  uint64_t full[kNSlot], empty[kNSlot];
  uint64_t tq_full[kTileQ], tq_empty[kTileQ];
  int4 sinfo[kNSlot];  // per stage: {first descriptor slot, list entries, ring offset of its bins, 0}
  int4 tq[kTileQ];  // {work item g (< 0: no more work), pool offset, list entries, 0}
};
static_assert(sizeof(PullCtl) <= kPullCtl, "control block");

// Consumer side of one tile: its n list entries arrive in stages of 1 .. kMaxGroup entries
// (as many as the producer could fit when it issued them) from stage `stage` on; one wait
// and one release per stage, not per entry.
template <typename T, int TW, int NV>
__device__ __forceinline__ void pull_consume_tile(const RoiFuseParams& p, const PullWs& ws,
                                                  const TileMap& tm, PullCtl& ctl,
                                                  const StageDesc* desc, const unsigned char* ring,
                                                  int block, int n, int& stage) {
  constexpr int V = VecOf<T>::n;
  constexpr int V2 = V / 2;
  constexpr int CG = 32 * V * NV;  // channels per group; NV > 1 requires C % CG == 0 (launcher)
  constexpr uint32_t kVecBytes = 32 * V * sizeof(T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = block / tm.groups, grp = block - t * tm.groups;
  const Tile tl = decode_tile(p, ws, tm, t, TW);
  const int C = p.C;
  const int c0 = grp * CG;
  const uint32_t bin_bytes = (uint32_t)min(CG, C - c0) * sizeof(T);
  const int y = tl.y0 + warp;
  const int cl = c0 + lane * V;
  const bool act = y <= tl.y1 && cl < C;
  uint64_t acc[TW][NV][V2];
#pragma unroll
  for (int x = 0; x < TW; ++x)
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int u = 0; u < V2; ++u) acc[x][v][u] = 0ull;

  for (int done = 0; done < n; ++stage) {
    const int slot = stage % kNSlot;
    mbar_wait(ctl.full + slot, (stage / kNSlot) & 1);
    const int4 info = ctl.sinfo[slot];
    uint32_t soff = (uint32_t)info.z;
    for (int ge = 0; ge < info.y; ++ge) {
      const StageDesc& d = desc[(info.x + ge) % kDescSlots];
      const int4 rd = d.rows[warp];
      const int npw = d.npw;
      if (act && rd.x >= 0 && !ARFE_SKIP(p, 1)) {  // (profiling aid: no math)
        const float a0 = __int_as_float(rd.z), a1 = __int_as_float(rd.w);
        const uint64_t a0p = pack2(a0, a0), a1p = pack2(a1, a1);
        const unsigned char* __restrict__ s0 =
            ring + soff + (uint32_t)(rd.x * npw) * bin_bytes + (uint32_t)(lane * V) * sizeof(T);
        const uint32_t two_step = rd.y ? (uint32_t)npw * bin_bytes : 0u;
#pragma unroll 2
        for (int jj = 0; jj < npw; ++jj) {
          uint64_t e[NV][V2];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            lds_pairs<T>(s0 + v * kVecBytes, e[v]);
#pragma unroll
            for (int u = 0; u < V2; ++u) e[v][u] = mul2(e[v][u], a0p);
          }
          if (rd.y) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              uint64_t e1[V2];
              lds_pairs<T>(s0 + two_step + v * kVecBytes, e1);
#pragma unroll
              for (int u = 0; u < V2; ++u) e[v][u] = fma2(e1[u], a1p, e[v][u]);
            }
          }
          float w[TW];
#pragma unroll
          for (int x = 0; x < TW; x += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(d.cw[jj] + x);
            w[x] = t4.x; w[x + 1] = t4.y; w[x + 2] = t4.z; w[x + 3] = t4.w;
          }
#pragma unroll
          for (int x = 0; x < TW; ++x) {
            const uint64_t wp = pack2(w[x], w[x]);
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
              for (int u = 0; u < V2; ++u) acc[x][v][u] = fma2(e[v][u], wp, acc[x][v][u]);
          }
          s0 += bin_bytes;
        }
      }
      soff += (uint32_t)(d.nph * npw) * bin_bytes;
    }
    done += info.y;
    __syncwarp();
    if (lane == 0) mbar_arrive(ctl.empty + slot);
  }
  // ---- every element of the tile written exactly once ----
  if (act) {
    float* __restrict__ dimg = p.dfeats[tl.l] + (size_t)tl.b * tl.H * tl.W * C;
#pragma unroll
    for (int x = 0; x < TW; ++x) {
      if (tl.x0 + x > tl.x1) continue;
      float* __restrict__ o = dimg + ((size_t)y * tl.W + tl.x0 + x) * C + cl;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int u = 0; u < V2; u += 2)
          *reinterpret_cast<ulonglong2*>(o + v * 32 * V + 2 * u) = make_ulonglong2(acc[x][v][u], acc[x][v][u + 1]);
    }
  }
}

// Work item g in [0, nb4 + nb8): (tile, channel group) of the narrow-tile pass,
// then of the wide-tile pass.
//
// ONE producer warp, vectorised: a pass of its loop takes the next list entries of the
// tile -- lane t does the bookkeeping and issues the bulk copies of the t-th of them (its
// 400-byte descriptor, its bin rows) -- and hands as many as fit the ring (up to kMaxGroup)
// to the consumers as ONE stage behind one mbarrier.  A single lane walking one entry at a
// time (~100 dependent instructions, ~0.5 us) used to be what the consumers waited for;
// one barrier round trip per entry was the next cost (profiles/r2_pull_stages.txt).
template <typename T, int NV, bool kNarrow>
__global__ void __launch_bounds__((kTileH + 1) * 32, (VecOf<T>::n * NV == 4 ? 3 : 2))
roi_bwd_pull_tma(const RoiFuseParams p, const PullWs ws, const TileMap tm4, const TileMap tm8,
                 int nb4, int nb8, int ring_bytes, int gmax) {
  constexpr int V = VecOf<T>::n;
  constexpr int CG = 32 * V * NV;
  extern __shared__ __align__(16) unsigned char smem[];
  PullCtl& ctl = *reinterpret_cast<PullCtl*>(smem);
  StageDesc* desc = reinterpret_cast<StageDesc*>(smem + kPullCtl);
  unsigned char* ring = smem + kPullCtl + kDescSlots * kDescBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = nb4 + nb8;

  if (tid == 0) {
    for (int i = 0; i < kNSlot; ++i) { mbar_init(ctl.full + i, 1); mbar_init(ctl.empty + i, kTileH); }
    for (int i = 0; i < kTileQ; ++i) { mbar_init(ctl.tq_full + i, 1); mbar_init(ctl.tq_empty + i, kTileH); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kTileH) {
    // ------------------------------------------------------------ producer
    const uint32_t ring_cap = (uint32_t)ring_bytes & ~127u;
    uint32_t head = 0, tail = 0;  // live bytes of the ring: [tail, head) modulo wrap
    uint32_t my_off = 0;          // lane s: ring offset of the stage in slot s
    int my_m = 0;                 // lane s: list entries of the stage in slot s
    int issued = 0, released = 0; // stages issued / known to be consumed
    int dissued = 0, dreleased = 0;  // descriptor slots handed out / free again
    const int C = p.C, BS = p.bin_stride;
    auto load_hdr = [&](int pool_off, int n, int first) -> int4 {
      return (n > 0 && first + lane < n) ? __ldg(reinterpret_cast<const int4*>(ws.pool + pool_off + first + lane))
                                         : make_int4(0, 0, 0, 0);
    };
    auto retire = [&]() {  // stage `released` was consumed: its ring space and descriptor slots are free
      dreleased += __shfl_sync(0xffffffffu, my_m, released % kNSlot);
      ++released;
      const uint32_t nxt = __shfl_sync(0xffffffffu, my_off, released % kNSlot);
      tail = released < issued ? nxt : head;
    };
    // the n list entries of work item g (pool entries from pool_off); hdr = headers of the first 32
    auto produce = [&](int g, int pool_off, int n, int4 hdr) {
      const bool narrow = g < nb4;
      const int grp = narrow ? g % tm4.groups : (g - nb4) % tm8.groups;
      const int c0 = grp * CG;
      const uint32_t bin_bytes = (uint32_t)min(CG, C - c0) * sizeof(T);
      const T* __restrict__ dsrc = static_cast<const T*>(p.dout) + c0;
      const StageDesc* __restrict__ gdesc = ws.pool + pool_off;
      int chunk0 = 0;  // hdr holds entries [chunk0, chunk0 + 32)
      int i = 0;       // next entry
      while (i < n) {
        if (i >= chunk0 + 32) {
          chunk0 = (i / 32) * 32;
          hdr = load_hdr(pool_off, n, chunk0);
        }
        const int it = i + lane;  // this lane's entry
        const bool valid = it < min(n, chunk0 + 32) && lane < gmax;
        const int hl = valid ? (it & 31) : 0;
        const int src_off = __shfl_sync(0xffffffffu, hdr.x, hl);
        const int nn = __shfl_sync(0xffffffffu, hdr.y, hl);
        const int rg = __shfl_sync(0xffffffffu, hdr.z, hl);
        const int nph = nn & 0xffff, npw = (nn >> 16) & 0xffff;
        const int nbins = nph * npw;
        const uint32_t bytes = valid ? (uint32_t)nbins * bin_bytes : 0u;
        uint32_t incl = bytes;  // inclusive prefix over the lanes
#pragma unroll
        for (int d = 1; d < kMaxGroup; d <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        const uint32_t bytes0 = __shfl_sync(0xffffffffu, bytes, 0);
        int m = 0;
        while (true) {
          while (released < issued && mbar_test(ctl.empty + released % kNSlot, (released / kNSlot) & 1)) retire();
          if (released == issued) head = tail = 0;  // ring empty
          const uint32_t room = head >= tail ? ring_cap - head : tail - head - 1u;
          // the entries of a tile are consecutive in the pool: their descriptors travel as ONE
          // copy, so a stage does not run across the end of the descriptor ring
          const int free_desc = min(kDescSlots - (dissued - dreleased), kDescSlots - dissued % kDescSlots);
          if (issued - released < kNSlot)
            m = __popc(__ballot_sync(0xffffffffu, valid && incl <= room && lane < free_desc));
          if (m > 0) break;
          if (issued - released < kNSlot && free_desc > 0 && head >= tail && released != issued && bytes0 < tail) {
            head = 0;  // wrap
            continue;
          }
          mbar_wait(ctl.empty + released % kNSlot, (released / kNSlot) & 1);  // wait for the oldest stage
          retire();
        }
        const int slot = issued % kNSlot;
        const uint32_t sbytes = __shfl_sync(0xffffffffu, incl, m - 1);
        if (lane == slot) { my_off = head; my_m = m; }
        if (lane == 0) {
          ctl.sinfo[slot] = make_int4(dissued % kDescSlots, m, (int)head, 0);
          mbar_arrive_expect_tx(ctl.full + slot, (ARFE_SKIP(p, 2) ? 0u : sbytes) + (uint32_t)m * kDescBytes);
        }
        __syncwarp();
        if (lane == 0) bulk_g2s(desc + dissued % kDescSlots, gdesc + i, (uint32_t)m * kDescBytes, ctl.full + slot);
        if (lane < m) {
          unsigned char* const dst = ring + head + incl - bytes;
          if (ARFE_SKIP(p, 2)) {
            // profiling aid: no bin copies
          } else if ((uint32_t)BS * sizeof(T) == bin_bytes) {
            // split layout, all channels in this group: the npw bins of a bin row are
            // contiguous -> one bulk copy per bin row
            for (int ih = 0; ih < nph; ++ih)
              bulk_g2s(dst + (uint32_t)(ih * npw) * bin_bytes,
                       dsrc + p.reg_off[rg] + (size_t)src_off + (size_t)(ih * p.PW) * BS,
                       (uint32_t)npw * bin_bytes, ctl.full + slot);
          } else {
            for (int bi = 0; bi < nbins; ++bi) {
              const int ih = bi / npw, iw = bi - ih * npw;
              bulk_g2s(dst + (uint32_t)bi * bin_bytes,
                       dsrc + p.reg_off[rg] + (size_t)src_off + (size_t)(ih * p.PW + iw) * BS, bin_bytes, ctl.full + slot);
            }
          }
        }
        __syncwarp();
        head += sbytes;
        ++issued;
        dissued += m;
        i += m;
      }
    };
    auto claim = [&]() -> int { return lane == 0 ? atomicAdd(ws.counters + 2, 1) : 0; };
    // Three dependent latencies precede a tile's first copy: the claim (an atomic), its tile
    // descriptor (a load that needs the claim) and its first 32 entry headers (a load that needs
    // the descriptor).  They are issued one loop iteration apart and each value is first touched
    // one iteration after its issue, so none of them is waited for while a tile is produced
    // (consumed in the iteration of their issue they stalled the producer ~1.7 us per tile).
    auto load_td = [&](int g) -> int2 {
      if (g >= total) return make_int2(0, -1);
      return g < nb4 ? __ldg(ws.tile_desc + tm4.tile_base + g / tm4.groups)
                     : __ldg(ws.tile_desc + tm8.tile_base + (g - nb4) / tm8.groups);
    };
    auto item = [&](int g, int2 td) -> int4 {
      // (profiling aid 4: every tile as if no region reached it -- the cost of the tile loop and the write-out)
      return make_int4(g < total ? g : -1, td.x, ARFE_SKIP(p, 4) ? min(td.y, 0) : td.y, 0);
    };
    int g0 = __shfl_sync(0xffffffffu, claim(), 0);
    int g1 = __shfl_sync(0xffffffffu, claim(), 0);
    int g2 = __shfl_sync(0xffffffffu, claim(), 0);
    int c3 = claim();
    int2 td0 = load_td(g0), td1 = load_td(g1), td2 = load_td(g2);
    int4 q0 = item(g0, td0), q1 = item(g1, td1);
    int4 h0 = load_hdr(q0.y, q0.z, 0), h1 = load_hdr(q1.y, q1.z, 0);
    for (int it = 0;; ++it) {
      const int g3 = __shfl_sync(0xffffffffu, c3, 0);   // claimed one iteration ago
      c3 = claim();
      const int2 td3 = load_td(g3);                      // touched next iteration
      const int4 q2 = item(g2, td2);                     // td2 was loaded one iteration ago
      const int4 h2 = load_hdr(q2.y, q2.z, 0);           // touched next iteration
      // announce the current item
      const int qs = it % kTileQ;
      if (it >= kTileQ) mbar_wait(ctl.tq_empty + qs, ((it / kTileQ) - 1) & 1);
      if (lane == 0) {
        ctl.tq[qs] = q0;
        mbar_arrive(ctl.tq_full + qs);
      }
      if (q0.x < 0) break;
      if (q0.z > 0) produce(q0.x, q0.y, q0.z, h0);  // < 0: inline tile, nothing to stream
      q0 = q1; h0 = h1;
      q1 = q2; h1 = h2;
      g2 = g3; td2 = td3;
    }
    return;
  }

  // ---------------------------------------------------------------- consumers
  int stage = 0;
  for (int it = 0;; ++it) {
    const int qs = it % kTileQ;
    mbar_wait(ctl.tq_full + qs, (it / kTileQ) & 1);
    const int4 q = ctl.tq[qs];
    __syncwarp();
    if (lane == 0) mbar_arrive(ctl.tq_empty + qs);  // copied to registers
    if (q.x < 0) break;
    const int n = q.z;
    if (n < 0) continue;  // served by roi_bwd_pull_inline
    if (kNarrow || q.x < nb4) pull_consume_tile<T, 4, NV>(p, ws, tm4, ctl, desc, ring, q.x, n, stage);
    else pull_consume_tile<T, 8, NV>(p, ws, tm8, ctl, desc, ring, q.x - nb4, n, stage);
  }
}

// ---- inline kernel: the tiles the binning kernel could not serve.  A small
// fixed grid walks the inline list; per tile and channel group: rounds of list
// building, expansion into shared memory and streaming from global memory.
template <typename T, int TW>
__device__ void pull_tile_inline(const RoiFuseParams& p, const PullWs& ws, const TileMap& tm, int t,
                                 ListSmem& sm, int4* rdesc, float* cwt) {
  constexpr int V = VecOf<T>::n;
  constexpr int V2 = V / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Tile tl = decode_tile(p, ws, tm, t, TW);
  const int C = p.C, BS = p.bin_stride, rowstep = p.PW * BS;
  const int y = tl.y0 + warp;
  for (int c0 = 0; c0 < C; c0 += 32 * V) {
    const int cl = c0 + lane * V;
    const bool act = y <= tl.y1 && cl < C;
    const T* __restrict__ dbase = static_cast<const T*>(p.dout) + cl;
    uint64_t acc[TW][V2];
#pragma unroll
    for (int x = 0; x < TW; ++x)
#pragma unroll
      for (int u = 0; u < V2; ++u) acc[x][u] = 0ull;
    __syncthreads();
    const int total_cand = init_candidates(sm, ws, tl.key);
    int pos = 0;
    while (pos < total_cand) {
      pos = build_list(sm, p, ws, tl, pos, total_cand, TW == 4 ? 4 : kJ);
      const int n = sm.list_n;
      for (int q0 = 0; q0 < n; q0 += kChunkR) {
        const int nc = min(kChunkR, n - q0);
        {  // thread == (tile row, entry)
          const int row = tid / kChunkR, q = tid - row * kChunkR;
          int4 d = make_int4(-1, 0, 0, 0);
          if (q < nc) {
            const ListEntry e = sm.list[q0 + q];
            d = expand_row(ws, e, tl.y0 + row);
            if (d.x >= 0) {  // -> {dout offset of (p0, pw0), npw | two << 8, w0, w1}
              const int k = e.src / p.R, r = e.src - k * p.R;
              d.x = (k * p.PH * p.PW + d.x * p.PW + e.pw0) * BS;
              d.y = e.npw | (d.y ? 256 : 0) | (r << 16);
            }
          }
          rdesc[row * kChunkR + q] = d;
        }
        {  // thread == (entry, pw)
          const int q = tid / kJ, jj = tid - q * kJ;
          if (q < nc) {
            const ListEntry e = sm.list[q0 + q];
            if (jj < e.npw) {
              float w[TW];
              expand_col<TW>(ws, e, jj, tl.x0, w);
#pragma unroll
              for (int x = 0; x < TW; x += 4)
                *reinterpret_cast<float4*>(cwt + tid * TW + x) = make_float4(w[x], w[x + 1], w[x + 2], w[x + 3]);
            }
          }
        }
        __syncthreads();
        if (act) {
          const int4* __restrict__ pr = rdesc + warp * kChunkR;
          for (int q = 0; q < nc; ++q) {
            const int4 cur = pr[q];
            if (cur.x < 0) continue;
            const int npw = cur.y & 255;
            const bool two = (cur.y & 256) != 0;
            const float a0 = __int_as_float(cur.z), a1 = __int_as_float(cur.w);
            const uint64_t a0p = pack2(a0, a0), a1p = pack2(a1, a1);
            const T* __restrict__ src = dbase + p.reg_off[(cur.y >> 16) & 3] + cur.x;
            const float* __restrict__ cw = cwt + q * (kJ * TW);
            for (int jj = 0; jj < npw; ++jj) {
              uint64_t e[V2], e1[V2];
              ldg_pairs<T>(src, e);
#pragma unroll
              for (int u = 0; u < V2; ++u) e[u] = mul2(e[u], a0p);
              if (two) {
                ldg_pairs<T>(src + rowstep, e1);
#pragma unroll
                for (int u = 0; u < V2; ++u) e[u] = fma2(e1[u], a1p, e[u]);
              }
#pragma unroll
              for (int x = 0; x < TW; ++x) {
                const uint64_t wp = pack2(cw[x], cw[x]);
#pragma unroll
                for (int u = 0; u < V2; ++u) acc[x][u] = fma2(e[u], wp, acc[x][u]);
              }
              src += BS;
              cw += TW;
            }
          }
        }
        __syncthreads();
      }
    }
    if (act) {
      float* __restrict__ dimg = p.dfeats[tl.l] + (size_t)tl.b * tl.H * tl.W * C;
#pragma unroll
      for (int x = 0; x < TW; ++x) {
        if (tl.x0 + x > tl.x1) continue;
        float* __restrict__ o = dimg + ((size_t)y * tl.W + tl.x0 + x) * C + cl;
#pragma unroll
        for (int u = 0; u < V2; u += 2)
          *reinterpret_cast<ulonglong2*>(o + 2 * u) = make_ulonglong2(acc[x][u], acc[x][u + 1]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
roi_bwd_pull_inline(const RoiFuseParams p, const PullWs ws, const TileMap tm4, const TileMap tm8, int nt4) {
  __shared__ ListSmem sm;
  __shared__ int4 rdesc[kTileH * kChunkR];
  __shared__ __align__(16) float cwt[kChunkR * kJ * kMaxTW];
  const int n = ws.counters[1];
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const int t = ws.inline_list[j];
    if (t < nt4) pull_tile_inline<T, 4>(p, ws, tm4, t, sm, rdesc, cwt);
    else pull_tile_inline<T, 8>(p, ws, tm8, t - nt4, sm, rdesc, cwt);
    __syncthreads();
  }
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_roi_fuse_forward_cl(const RoiFuseParams& p, int dtype, int out_cl,
                                       cudaStream_t stream) {
  const int PHW = p.PH * p.PW;
  int grid = p.K * p.R;
  if (p.flag_list && grid > 296) grid = 296;  // behind the ring kernel: usually nothing to do
  cudaError_t e;
  int opitch = 0, bpp = PHW, smem = kHdrBytes;
  if (!out_cl) {
    opitch = p.C + 4;
    bpp = (96 * 1024) / (opitch * 4);
    if (bpp < 1) return cudaErrorInvalidValue;
    if (bpp > PHW) bpp = PHW;
    smem += bpp * opitch * 4;
  }
  static const int occ = ARFE_KNOB_ENV("ARFE_FWD_OCC", 3);
#define ARFE_FWD_CL1(TT, OC, OCC)                                                       \
  do {                                                                                  \
    if ((e = set_smem(roi_fuse_fwd_cl<TT, OC, OCC>, smem)) != cudaSuccess) return e;    \
    roi_fuse_fwd_cl<TT, OC, OCC><<<grid, kThreads, smem, stream>>>(p, opitch, bpp);     \
  } while (0)
#define ARFE_FWD_CL(TT, OC)                                                             \
  do {                                                                                  \
    if (occ == 3) ARFE_FWD_CL1(TT, OC, 3); else ARFE_FWD_CL1(TT, OC, 4);                \
  } while (0)
  if (dtype == 0) { if (out_cl) ARFE_FWD_CL(float, true); else ARFE_FWD_CL(float, false); }
  else { if (out_cl) ARFE_FWD_CL(__nv_bfloat16, true); else ARFE_FWD_CL(__nv_bfloat16, false); }
#undef ARFE_FWD_CL1
#undef ARFE_FWD_CL
  return cudaGetLastError();
}

static cudaError_t launch_prep(const RoiFuseParams& p, const PullWs& ws, cudaStream_t stream) {
  const int N = p.K * p.R;
  const int prep_smem = ws.nkeys * (kPrepBlock / 32) * 4;
  if (prep_smem > 160 * 1024 || ws.nblk > kMaxPrepBlocks) return cudaErrorInvalidValue;
  cudaError_t e = set_smem(roi_prep_kernel, prep_smem);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ws.counters + 3, 0, 12, stream)) != cudaSuccess) return e;
  roi_prep_kernel<<<ws.nblk + N, kPrepThreads, prep_smem, stream>>>(p, ws);
  return cudaGetLastError();
}

// Forward with a plan: roi_prep_kernel writes the region tables into the
// workspace (the backward can reuse them: plan_ready), the ring kernel consumes
// them, the L1-cached kernel serves the regions the ring kernel left out.
// Channels-last features and output.
// stages: 1 = build the plan (roi_prep_kernel), 2 = run the forward on it.
cudaError_t launch_roi_fuse_forward_plan(const RoiFuseParams& p0, int dtype, void* workspace,
                                         size_t workspace_bytes, int stages, cudaStream_t stream) {
  RoiFuseParams p = p0;
  const int N = p.K * p.R;
  PullWs ws;
  const size_t need = pull_ws_layout(N, p.L, p.B, p.H, p.W, static_cast<unsigned char*>(workspace), &ws);
  if (need > workspace_bytes) return cudaErrorInvalidValue;
  const int V = dtype == 0 ? 4 : 8, elt = dtype == 0 ? 4 : 2;
  // consumer warp == (output column, channel chunk pair | chunk)
  static const int nch_env = ARFE_KNOB_ENV("ARFE_FWD_NCH", 0);
  const int nch = (p.C % (64 * V) == 0 && p.PH * V * 2 <= 64 && nch_env != 1) ? 2 : 1;
  const int ncons = p.PW * ((p.C + 32 * V * nch - 1) / (32 * V * nch));
  const bool ring_ok = (p.PH == 7 || p.PH == 14) && p.PH * V <= 64 && ncons <= (nch == 2 ? 7 : 14);
  const int threads = (ncons + 1) * 32;
  // two CTAs per SM when the block is small (<= 256 threads at 128 registers) or
  // the accumulators are few (7x7 fp32 single chunk at 64 registers), else one
  const int per_sm = (threads <= 256 || p.PH * V * nch <= 32) ? 2 : 1;
  const int fixed = kFwdTabs * (int)sizeof(FwdTab) + 512;
  const int ring = (per_sm == 2 ? 108 * 1024 : 200 * 1024) - fixed;
  ws.fwd_wlen_cap = ring_ok ? ring / (p.C * elt) : 0;  // a window row must fit the ring
  cudaError_t e = cudaSuccess;
  if ((stages & 1) && (e = launch_prep(p, ws, stream)) != cudaSuccess) return e;
  if (!(stages & 2)) return cudaSuccess;
  if (!ring_ok) return launch_roi_fuse_forward_cl(p, dtype, 1, stream);
  // plan built earlier (plan_ready): the ring kernel's work counter is the only state a forward
  // consumes, so the same plan serves any number of forward calls
  if (!(stages & 1) && (e = cudaMemsetAsync(ws.counters + 5, 0, 4, stream)) != cudaSuccess) return e;
  const int sms = sm_count();
  const int pgrid = N < per_sm * sms ? N : per_sm * sms;
  const int smem = fixed + ring;
#define ARFE_FWD_RING(TT, PHH, NCHH, NTT)                                                             \
  do {                                                                                                \
    if ((e = set_smem(roi_fuse_fwd_ring<TT, PHH, NCHH, NTT>, smem)) != cudaSuccess) return e;         \
    roi_fuse_fwd_ring<TT, PHH, NCHH, NTT><<<pgrid, threads, smem, stream>>>(p, ws, ncons, ring);      \
  } while (0)
  if (dtype == 0) {
    if (p.PH == 7) {
      if (nch == 2) ARFE_FWD_RING(float, 7, 2, 256);
      else if (threads <= 256) ARFE_FWD_RING(float, 7, 1, 256);
      else ARFE_FWD_RING(float, 7, 1, 480);
    } else {
      if (threads <= 256) ARFE_FWD_RING(float, 14, 1, 256); else ARFE_FWD_RING(float, 14, 1, 480);
    }
  } else {
    if (threads <= 256) ARFE_FWD_RING(__nv_bfloat16, 7, 1, 256); else ARFE_FWD_RING(__nv_bfloat16, 7, 1, 480);
  }
#undef ARFE_FWD_RING
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  p.flag_list = ws.fwd_list;
  p.flag_count = ws.counters + 4;
  return launch_roi_fuse_forward_cl(p, dtype, 1, stream);
}

size_t roi_pull_workspace_bytes(int K, int R, int L, int B, const int* H, const int* W) {
  return pull_ws_layout(K * R, L, B, H, W, nullptr, nullptr);
}

// dout: channels-last [K][PH*PW][R*C]; dfeats: NHWC fp32, fully written.
// Regions whose tables did not fit are flagged in the workspace and added
// afterwards by the atomic kernel.
// stages: 1 = build the plan, 2 = bin the tiles (stage descriptors), 4 = pull.
cudaError_t launch_roi_fuse_backward_pull(const RoiFuseParams& p, int dtype, void* workspace,
                                          size_t workspace_bytes, int stages, cudaStream_t stream) {
  const int N = p.K * p.R;
  PullWs ws;
  const size_t need = pull_ws_layout(N, p.L, p.B, p.H, p.W, static_cast<unsigned char*>(workspace), &ws);
  if (need > workspace_bytes) return cudaErrorInvalidValue;
  if (ws.nblk > kMaxPrepBlocks) return cudaErrorInvalidValue;
  if ((long long)p.K * p.PH * p.PW * p.R * p.C > 0x7fffffffLL) return cudaErrorInvalidValue;  // 32-bit dout offsets
  cudaError_t e;
  if ((stages & 1) && (e = launch_prep(p, ws, stream)) != cudaSuccess) return e;
  if ((stages & 2) && (e = cudaMemsetAsync(ws.counters, 0, 12, stream)) != cudaSuccess) return e;
  // bins built earlier (arfe_roi_pull_bin): only the pull kernel's work counter is reset, so the
  // same bins serve any number of backward calls
  if (stages == 4 && (e = cudaMemsetAsync(ws.counters + 2, 0, 4, stream)) != cudaSuccess) return e;
  // Two tile shapes: the small upper-level maps carry ~40x more region-pixels per
  // tile than level 0, so they get narrow tiles and go first (heaviest level
  // first inside each launch); the big maps follow with wide tiles.
  const int V = dtype == 0 ? 4 : 8;
  // fp32: a CTA takes 2 x 128 channels when C allows (1 KB bin pieces: half the bulk
  // copies and stages of the 128-channel variant)
  static const int nv_env = ARFE_KNOB_ENV("ARFE_PULL_NV", 0);
  const int nv = (dtype == 0 && p.C % (64 * V) == 0 && nv_env != 1) ? 2 : 1;
  TileMap tm[2];
  int ntiles[2];
  int tile_base = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int tw = pass == 0 ? 4 : 8;
    TileMap& m = tm[pass];
    int total = 0, ns = 0;
    m.groups = (p.C + 32 * V * nv - 1) / (32 * V * nv);
    for (int l = p.L - 1; l >= 0; --l) {
      if (pull_heavy(p.H[l], p.W[l]) != (pass == 0)) continue;
      m.tstart[ns] = total;
      m.level[ns] = l;
      m.tiles_x[ns] = (p.W[l] + tw - 1) / tw;
      m.tiles_y[ns] = (p.H[l] + kTileH - 1) / kTileH;
      total += m.tiles_x[ns] * m.tiles_y[ns] * p.B;
      ++ns;
    }
    m.nslots = ns;
    for (int j2 = ns; j2 <= kMaxLevels; ++j2) m.tstart[j2] = total;
    for (int j2 = ns; j2 < kMaxLevels; ++j2) { m.level[j2] = 0; m.tiles_x[j2] = m.tiles_y[j2] = 1; }
    m.tile_base = tile_base;
    tile_base += total;
    ntiles[pass] = total;
  }
  if (ntiles[0] + ntiles[1] == 0) return cudaSuccess;
  if (stages & 2) {
    roi_bin_kernel<<<ntiles[0] + ntiles[1], kBinThreads, 0, stream>>>(p, ws, tm[0], tm[1], ntiles[0]);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if (!(stages & 4)) return cudaSuccess;
  const bool narrow = ntiles[1] == 0;  // every tile is 4 pixels wide
  static const int persm_env = ARFE_KNOB_ENV("ARFE_PULL_PERSM", 0);
  const int per_sm = persm_env ? persm_env : ((dtype == 0 && nv == 1) ? 3 : 2);
  static const int gmax_env = ARFE_KNOB_ENV("ARFE_PULL_G", kMaxGroup);
  const int gmax = gmax_env < 1 ? 1 : (gmax_env > kMaxGroup ? kMaxGroup : gmax_env);
  const int ring = (((per_sm == 3 ? 75 : 113) * 1024 - kPullCtl - kDescSlots * kDescBytes) / 1024) * 1024;
  const int jmax = narrow ? 4 : kJ;
  const int max_entry = (p.PH < kMaxPh ? p.PH : kMaxPh) * (p.PW < jmax ? p.PW : jmax) * 32 * V * nv * (dtype == 0 ? 4 : 2);
  if (max_entry > ring) return cudaErrorInvalidValue;  // (C <= 512 per group by construction: cannot happen)
  const int smem = kPullCtl + kDescSlots * kDescBytes + ring;
  const int nb4 = ntiles[0] * tm[0].groups, nb8 = ntiles[1] * tm[1].groups;
  const int sms = sm_count();
  const int pgrid = nb4 + nb8 < per_sm * sms ? nb4 + nb8 : per_sm * sms;
#define ARFE_PULL(TT, NVV, NAR)                                                                     \
  do {                                                                                              \
    if ((e = set_smem(roi_bwd_pull_tma<TT, NVV, NAR>, smem)) != cudaSuccess) return e;              \
    roi_bwd_pull_tma<TT, NVV, NAR><<<pgrid, (kTileH + 1) * 32, smem, stream>>>(p, ws, tm[0], tm[1], nb4, nb8, ring, gmax); \
  } while (0)
#define ARFE_PULL_N(TT, NVV) do { if (narrow) ARFE_PULL(TT, NVV, true); else ARFE_PULL(TT, NVV, false); } while (0)
  if (dtype == 0) {
    if (nv == 2) ARFE_PULL_N(float, 2);
    else ARFE_PULL_N(float, 1);
  } else ARFE_PULL_N(__nv_bfloat16, 1);
#undef ARFE_PULL_N
#undef ARFE_PULL
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const int igrid = ntiles[0] + ntiles[1] < 592 ? ntiles[0] + ntiles[1] : 592;
  if (dtype == 0) roi_bwd_pull_inline<float><<<igrid, kThreads, 0, stream>>>(p, ws, tm[0], tm[1], ntiles[0]);
  else roi_bwd_pull_inline<__nv_bfloat16><<<igrid, kThreads, 0, stream>>>(p, ws, tm[0], tm[1], ntiles[0]);
  return cudaGetLastError();
}

// Device pointer to the list of regions whose tables did not fit (the atomic
// fallback kernel walks it).
const int* roi_pull_flag_list(int K, int R, int L, int B, const int* H, const int* W, void* workspace,
                              const int** count) {
  PullWs ws;
  pull_ws_layout(K * R, L, B, H, W, static_cast<unsigned char*>(workspace), &ws);
  *count = ws.counters + 3;
  return ws.flag_list;
}

}  // namespace arfe
