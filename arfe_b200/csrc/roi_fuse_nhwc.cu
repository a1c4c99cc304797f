// Channels-last (NHWC) fast path of the AR-RFF extraction.
//
// With channels innermost every access of this gather/scatter workload is
// lane == channel coalesced, so nothing has to be transposed through shared
// memory and nothing needs an atomic:
//
//   forward  : one CTA per (RoI, region); a warp owns a bin, lanes own 4 (fp32)
//              or 8 (bf16) consecutive channels; every tap is ONE 128-bit
//              L1-cached load per lane (100 % sector efficiency; taps shared by
//              neighbouring bins hit L1).  Output either channels-last
//              (direct 128-bit stores) or NCHW (staged [bin][C] in smem).
//   backward : PULL formulation.  roi_prep writes, per region, its window and
//              for every window row/column the bins that sample it with their
//              aggregated weights (the transpose of the forward tables) into a
//              caller-provided workspace.  roi_bwd_pull then gives every 8x8
//              pixel tile of every level to one CTA: warp == tile row,
//              lanes == channels, accumulators in registers, the regions that
//              intersect the tile are visited in index order -> every gradient
//              element is written exactly once, deterministically, with no
//              atomics and no zero-fill.  (The NCHW/atomic kernel is bound by
//              the L2 atomic units at ~1 fp32 element per slice-clock; see
//              DESIGN.md section 6.)
#include <stdlib.h>

#include "roi_common.cuh"

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

// dynamic smem: [CtaHeader][AxisTable y][AxisTable x][outs: bins_per_pass * opitch] (NCHW out only)
template <typename T, bool kOutCL, int kOcc>
__global__ void __launch_bounds__(kThreads, kOcc)
roi_fuse_fwd_cl(const RoiFuseParams p, int opitch, int bins_per_pass) {
  constexpr int V = VecOf<T>::n;
  extern __shared__ __align__(16) unsigned char smem[];
  CtaHeader& hd = *reinterpret_cast<CtaHeader*>(smem);
  AxisTable& ty = *reinterpret_cast<AxisTable*>(smem + 128);
  AxisTable& tx = *reinterpret_cast<AxisTable*>(smem + 128 + sizeof(AxisTable));
  float* outs = reinterpret_cast<float*>(smem + kHdrBytes);

  const int k = blockIdx.x / p.R, r = blockIdx.x % p.R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PH = p.PH, PW = p.PW, PHW = PH * PW, C = p.C, RC = p.R * C;
  T* __restrict__ out = static_cast<T*>(p.out);
  // element (bin, c) of this region's output block
  auto out_index = [&](int bin, int c) -> size_t {
    return kOutCL ? ((size_t)k * PHW + bin) * RC + (size_t)r * C + c
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
    // warp == (bin row ph, channel chunk): the warps of neighbouring bin rows run
    // side by side and walk their rows in opposite directions (even ph top-down,
    // odd ph bottom-up), so the feature row two bin rows share is touched by
    // both at about the same time and the second touch hits L1.
    const int nchunk = (C + cw - 1) / cw;
    for (int item = warp; item < PH * nchunk; item += kWarps) {
      const int chunk = item / PH, ph = item - chunk * PH;
      const int c = chunk * cw + lane * V;
      if (c >= C) continue;
      const int nr = ty.cnt[ph];
      const float* __restrict__ wy = ty.w + ty.off[ph];
      const bool up = (ph & 1) != 0;
      const T* __restrict__ rowbase = fimg + (size_t)ty.first[ph] * rowstride + c;
      T* __restrict__ o = out + out_index(ph * PW, c);
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
        st_vec<T>(o + (size_t)pw * RC, f);
      }
    }
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

// ------------------------------------------------------------ region prep
// Workspace layout (bytes, each array 256-byte aligned):
//   hdr     : N * 32                         RegionHdr
//   rowtab  : N * kTabLen * 16               TapEntry per (bin block, window row)   (transposed table)
//   colbin  : N * kMaxPool * 8               ColBin per output column pw            (forward table)
//   colw    : N * kTabLen * 4                aggregated column weights, bin-major
//   seg_ids : nblk * NK * kPrepBlock * 4     region ids per (prep block, key), index order;
//   seg_cnt : nblk * NK * 4                  key = (level, image, 8-row band of the level)
constexpr int kTabLen = 256;    // per region: (row bin blocks) x (window rows) <= kTabLen, column weights <= kTabLen
constexpr int kMaxBlk = 15;     // a row may be sampled by up to 2 * kMaxBlk bins (4-bit field)
constexpr int kPrepBlock = 256; // regions per header block

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

struct PullWs {
  RegionHdr* hdr;
  TapEntry* rowtab;
  ColBin* colbin;
  float* colw;
  int* seg_ids;
  int* seg_cnt;
  int nblk;
  int nkeys;                  // sum over levels of B * bands(level)
  int key_base[kMaxLevels];   // first key of level l
  int nbands[kMaxLevels];     // ceil(H_l / 8)
};

constexpr int kBandH = 8;     // rows per band == rows per pull tile

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline size_t pull_ws_layout(int N, int L, int B, const int* H, unsigned char* base, PullWs* ws) {
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
  if (ws) {
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

// Blocks [0, nblk): headers of kPrepBlock regions each + per-level ordered id
// segments.  Blocks [nblk, nblk + N): tap tables of one region each.
__global__ void __launch_bounds__(kPrepBlock)
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
    for (int e = tid; e < ws.nkeys * kW; e += kPrepBlock) mask[e] = 0u;
    __syncthreads();
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
    for (int key = k0; key <= k1; ++key) atomicOr(&mask[key * kW + (tid >> 5)], 1u << (tid & 31));
    __syncthreads();
    for (int key = k0; key <= k1; ++key) {
      int pos = __popc(mask[key * kW + (tid >> 5)] & ((1u << (tid & 31)) - 1u));
      for (int w = 0; w < (tid >> 5); ++w) pos += __popc(mask[key * kW + w]);
      ws.seg_ids[((size_t)blockIdx.x * ws.nkeys + key) * kPrepBlock + pos] = i;
    }
    for (int key = tid; key < ws.nkeys; key += kPrepBlock) {
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
      for (int j = t2; j < ctotal; j += kPrepBlock - 32) ws.colw[(size_t)i * kTabLen + j] = tx.w[j];
    }
    __syncthreads();
    if (!fit || ctotal > kTabLen) h.flags = 1;
    else h.flags = fit << 8;  // row bin blocks
  }
  if (tid == 0) ws.hdr[i] = h;
}

// ----------------------------------------------------------- pull backward
constexpr int kTileH = kBandH;  // tile rows, warp == tile row
constexpr int kListCap = 512;   // list entries per pass
constexpr int kChunkR = 32;     // list entries expanded per round
constexpr int kJ = 8;           // output columns (pw) per list entry
constexpr int kMaxPrepBlocks = 511;   // K * regions <= 130 816 for the pull path
static_assert(kTileH * kChunkR == kThreads && kChunkR * kJ == kThreads, "expand: one item per thread");

struct TileMap {
  int start[kMaxLevels + 1];  // first CTA of each scheduled slot
  int level[kMaxLevels];      // level of slot j (heaviest first)
  int tiles_x[kMaxLevels], tiles_y[kMaxLevels];
  int nslots;
  int groups;                 // channel groups per tile (C / (32 * V * NV))
};

// Compact copy of a listed (region, row bin block, pw block) in shared memory.
struct __align__(16) ListEntry {
  int id;            // region id | row block << 24
  int src;           // k * R + r
  short ymin, ymax;  // window rows
  short pw0, npw;    // output columns [pw0, pw0 + npw) reach this tile's columns
};

// NV = 128-bit vectors per lane, TW = tile width in pixels.
//
// The bilinear sum is separable: a gradient pixel (y, x) of a region receives
//     sum_ph ry[ph][y] * sum_pw cx[pw][x] * dout[ph][pw]
// with ry / cx the aggregated row / column weights.  A warp owns one tile row
// y, lanes own channels.  For one listed region the warp walks the few output
// columns pw whose samples reach the tile's columns; for each it loads
// dout[ph][pw] of the (<= 2) bins ph sampling row y ONCE, combines them with the
// row weights, and adds the result to its TW pixels with that pw's column
// weights (zero where pw does not reach) -- 2 loads per pw instead of 4 per pixel.
//
// Round structure (kChunkR list entries at a time, one item per thread):
//   expand : thread (tile row, entry): the row's TapEntry -> {dout offset, bins,
//            row weights}; thread (entry, pw): that pw's TW column weights.
//   stream : warp == tile row, lanes == channels, as above, FFMA2 arithmetic.
template <typename T, int NV, int TW>
__global__ void __launch_bounds__(kThreads, (NV * TW * VecOf<T>::n <= 16 ? 4 : (NV * TW * VecOf<T>::n <= 32 ? 3 : 2)))
roi_bwd_pull(const RoiFuseParams p, const PullWs ws, const TileMap tm) {
  constexpr int V = VecOf<T>::n;
  constexpr int V2 = V / 2;
  __shared__ ListEntry list[kListCap];
  __shared__ int4 rdesc[kTileH * kChunkR + 1];             // {dout offset (<0: none), npw | two << 8, w0, w1}
  __shared__ __align__(16) float cwt[kChunkR * kJ * TW];   // [entry][pw - pw0][tile column]
  __shared__ int list_n;
  __shared__ int warp_tot[kThreads / 32];
  __shared__ int pre[kMaxPrepBlocks + 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int j = 0;
  while (j + 1 < tm.nslots && (int)blockIdx.x >= tm.start[j + 1]) ++j;
  const int l = tm.level[j];
  int t = blockIdx.x - tm.start[j];
  const int grp = t % tm.groups;
  t /= tm.groups;
  const int per_img = tm.tiles_x[j] * tm.tiles_y[j];
  const int b = t / per_img;
  t -= b * per_img;
  const int tyi = t / tm.tiles_x[j], txi = t - tyi * tm.tiles_x[j];
  const int y0 = tyi * kTileH, x0 = txi * TW;
  const int key = ws.key_base[l] + b * ws.nbands[l] + tyi;  // this tile's (level, image, band)
  const int H = p.H[l], W = p.W[l], C = p.C, RC = p.R * C, PHW = p.PH * p.PW, PW = p.PW;
  const int y = y0 + warp;  // this warp's row
  const int y1 = min(y0 + kTileH, H) - 1, x1 = min(x0 + TW, W) - 1;
  const int cl = grp * (32 * V * NV) + lane * V;  // first channel of this lane's first vector
  const T* __restrict__ dbase = static_cast<const T*>(p.dout) + cl;
  float* __restrict__ dimg = p.dfeats[l] + (size_t)b * H * W * C;
  const int rowstep = PW * RC;
  const bool cok = cl < C;  // NV > 1 requires C % (32 * V * NV) == 0 (launcher)

  uint64_t acc[TW][NV][V2];
#pragma unroll
  for (int x = 0; x < TW; ++x)
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int u = 0; u < V2; ++u) acc[x][v][u] = 0ull;

  // candidates = the ids of this tile's (level, image, band) key over all prep
  // blocks, block-major == region-index order; pre[] = exclusive prefix of the
  // per-block counts so that one 256-thread batch spans blocks
  for (int i = tid; i <= ws.nblk; i += kThreads)
    pre[i] = i < ws.nblk ? ws.seg_cnt[(size_t)i * ws.nkeys + key] : 0;
  if (tid == 0) rdesc[kTileH * kChunkR] = make_int4(-1, 0, 0, 0);  // prefetch pad
  __syncthreads();
  if (warp == 0) {  // exclusive scan of pre[0..nblk] by one warp
    int carry = 0;
    for (int i0 = 0; i0 <= ws.nblk; i0 += 32) {
      const int i = i0 + lane;
      const int v = i <= ws.nblk ? pre[i] : 0;
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t2 = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t2;
      }
      if (i <= ws.nblk) pre[i] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  const int total_cand = pre[ws.nblk];

  int pos = 0;  // cursor into the flattened candidates (uniform)
  bool more = true;
  while (more) {
    __syncthreads();
    if (tid == 0) list_n = 0;
    __syncthreads();
    // ---- ordered compaction of the intersecting regions into list[] ----
    bool full = false;
    while (pos < total_cand && !full) {
      const int q = pos + tid;
      int mine = 0, nbr = 0, ncb = 0, id = 0, plo = 0, phi = -1;
      RegionHdr h;
      if (q < total_cand) {
        int lo = 0, hi = ws.nblk - 1;  // last block with pre[blk] <= q
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (pre[mid] <= q) lo = mid; else hi = mid - 1;
        }
        id = ws.seg_ids[((size_t)lo * ws.nkeys + key) * kPrepBlock + (q - pre[lo])];
        h = ws.hdr[id];
        const bool hit = h.lvl == l && h.batch == b && (h.flags & 1) == 0 && h.ymax >= y0 &&
                         h.ymin <= y1 && h.xmax >= x0 && h.xmin <= x1;
        if (hit) {
          // output columns whose samples reach [x0, x1]
          const ColBin* __restrict__ cbp = ws.colbin + (size_t)id * kMaxPool;
          plo = PW;
          for (int pw = 0; pw < PW; ++pw) {
            const ColBin c = cbp[pw];
            if (c.cnt > 0 && c.first <= x1 && c.first + c.cnt - 1 >= x0) {
              plo = min(plo, pw);
              phi = pw;
            }
          }
          if (phi >= 0) {
            nbr = (h.flags >> 8) & 15;
            ncb = (phi - plo + kJ) / kJ;
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
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      int base = list_n, tot = 0;
      for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) base += warp_tot[w];
        tot += warp_tot[w];
      }
      if (list_n + tot > kListCap) { full = true; __syncthreads(); break; }
      if (mine) {
        ListEntry e;
        e.src = h.src;
        e.ymin = (short)h.ymin; e.ymax = (short)h.ymax;
        int o = base + incl - mine;
        for (int rb = 0; rb < nbr; ++rb)
          for (int cb = 0; cb < ncb; ++cb) {
            e.id = id | (rb << 24);
            e.pw0 = (short)(plo + cb * kJ);
            e.npw = (short)min(kJ, phi - (plo + cb * kJ) + 1);
            list[o++] = e;
          }
      }
      __syncthreads();
      if (tid == 0) list_n += tot;
      __syncthreads();
      pos += kThreads;
    }
    more = pos < total_cand;
    __syncthreads();
    const int n = list_n;

    for (int c0 = 0; c0 < n; c0 += kChunkR) {
      const int nc = min(kChunkR, n - c0);
      // ---- expand, one item per thread ----
      {  // thread == (tile row, entry): row descriptor
        const int row = tid / kChunkR, q = tid - row * kChunkR;
        int4 d = make_int4(-1, 0, 0, 0);
        if (q < nc) {
          const ListEntry e = list[c0 + q];
          const int yy = y0 + row;
          if (yy >= e.ymin && yy <= e.ymax) {
            const int id = e.id & 0xffffff, rb = e.id >> 24;
            const TapEntry re = ws.rowtab[(size_t)id * kTabLen + tab_index(rb, yy - e.ymin, e.ymax - e.ymin + 1)];
            if (re.n > 0) {
              const int k = e.src / p.R, r = e.src - k * p.R;
              d.x = (k * PHW + re.p0 * PW + e.pw0) * RC + r * C;
              d.y = e.npw | (re.n > 1 ? 256 : 0);
              d.z = __float_as_int(re.w0);
              d.w = __float_as_int(re.w1);
            }
          }
        }
        rdesc[row * kChunkR + q] = d;
      }
      {  // thread == (entry, pw): the pw's column weights over the tile
        const int q = tid / kJ, jj = tid - q * kJ;
        if (q < nc) {
          const ListEntry e = list[c0 + q];
          if (jj < e.npw) {
            const int id = e.id & 0xffffff;
            const ColBin c = ws.colbin[(size_t)id * kMaxPool + e.pw0 + jj];
            const float* __restrict__ cwp = ws.colw + (size_t)id * kTabLen + c.off;
            float w[TW];
#pragma unroll
            for (int x = 0; x < TW; ++x) {
              const int i = x0 + x - c.first;
              w[x] = (i >= 0 && i < c.cnt) ? __ldg(cwp + i) : 0.f;
            }
#pragma unroll
            for (int x = 0; x < TW; x += 4)
              *reinterpret_cast<float4*>(cwt + tid * TW + x) = make_float4(w[x], w[x + 1], w[x + 2], w[x + 3]);
          }
        }
      }
      __syncthreads();
      // ---- stream: warp == row y, lanes == channels ----
      if (y <= y1 && cok) {
        const int4* __restrict__ pr = rdesc + warp * kChunkR;
        int4 rd = pr[0];
        for (int q = 0; q < nc; ++q) {
          const int4 cur = rd;
          rd = pr[q + 1];  // prefetch (the array is padded by one)
          if (cur.x < 0) continue;
          const int npw = cur.y & 255;
          const bool two = (cur.y & 256) != 0;
          const float a0 = __int_as_float(cur.z), a1 = __int_as_float(cur.w);
          const uint64_t a0p = pack2(a0, a0), a1p = pack2(a1, a1);
          const T* __restrict__ src = dbase + cur.x;
          const float* __restrict__ cw = cwt + q * (kJ * TW);
#pragma unroll 2
          for (int jj = 0; jj < npw; ++jj) {
            uint64_t e[NV][V2];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              uint64_t d0[V2];
              ldg_pairs<T>(src + v * 32 * V, d0);
#pragma unroll
              for (int u = 0; u < V2; ++u) e[v][u] = mul2(d0[u], a0p);
            }
            if (two) {
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                uint64_t d1[V2];
                ldg_pairs<T>(src + rowstep + v * 32 * V, d1);
#pragma unroll
                for (int u = 0; u < V2; ++u) e[v][u] = fma2(d1[u], a1p, e[v][u]);
              }
            }
            float w[TW];
#pragma unroll
            for (int x = 0; x < TW; x += 4) {
              const float4 t4 = *reinterpret_cast<const float4*>(cw + x);
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
            src += RC;
            cw += TW;
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- every element of the tile written exactly once ----
  if (y <= y1 && cok) {
#pragma unroll
    for (int x = 0; x < TW; ++x) {
      if (x0 + x > x1) continue;
      float* __restrict__ o = dimg + ((size_t)y * W + x0 + x) * C + cl;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int u = 0; u < V2; u += 2)
          *reinterpret_cast<ulonglong2*>(o + v * 32 * V + 2 * u) = make_ulonglong2(acc[x][v][u], acc[x][v][u + 1]);
      }
    }
  }
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_roi_fuse_forward_cl(const RoiFuseParams& p, int dtype, int out_cl,
                                       cudaStream_t stream) {
  const int PHW = p.PH * p.PW;
  const int grid = p.K * p.R;
  int opitch = 0, bpp = PHW, smem = kHdrBytes;
  if (!out_cl) {
    opitch = p.C + 4;
    bpp = (96 * 1024) / (opitch * 4);
    if (bpp < 1) return cudaErrorInvalidValue;
    if (bpp > PHW) bpp = PHW;
    smem += bpp * opitch * 4;
  }
  cudaError_t e;
  static const int occ = [] { const char* e = getenv("ARFE_FWD_OCC"); return e ? atoi(e) : 4; }();
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

size_t roi_pull_workspace_bytes(int K, int R, int L, int B, const int* H) {
  return pull_ws_layout(K * R, L, B, H, nullptr, nullptr);
}

// dout: channels-last [K][PH*PW][R*C]; dfeats: NHWC fp32, fully written.
// `flags_out` (host side) is not needed: regions whose tables did not fit are
// flagged in the workspace and added afterwards by the atomic kernel.
cudaError_t launch_roi_fuse_backward_pull(const RoiFuseParams& p, int dtype, void* workspace,
                                          size_t workspace_bytes, cudaStream_t stream) {
  const int N = p.K * p.R;
  PullWs ws;
  const size_t need = pull_ws_layout(N, p.L, p.B, p.H, static_cast<unsigned char*>(workspace), &ws);
  if (need > workspace_bytes) return cudaErrorInvalidValue;
  const int prep_smem = ws.nkeys * (kPrepBlock / 32) * 4;
  if (prep_smem > 160 * 1024 || ws.nblk > kMaxPrepBlocks) return cudaErrorInvalidValue;
  cudaError_t e = set_smem(roi_prep_kernel, prep_smem);
  if (e != cudaSuccess) return e;
  roi_prep_kernel<<<ws.nblk + N, kPrepBlock, prep_smem, stream>>>(p, ws);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // Two launches: the small upper-level maps carry ~40x more region-pixels per
  // tile than level 0, so they get narrow tiles + a channel split and go first
  // (heaviest level first inside each launch); the big maps follow with wide tiles.
  const int V = dtype == 0 ? 4 : 8;
  if ((long long)p.K * p.PH * p.PW * p.R * p.C > 0x7fffffffLL) return cudaErrorInvalidValue;  // 32-bit dout offsets
  auto heavy = [&](int l) { return (long long)p.H[l] * p.W[l] <= 64 * 96; };
  for (int pass = 0; pass < 2; ++pass) {
    const int tw = pass == 0 ? 4 : 8;
    const int nv = (pass == 0 || dtype != 0) ? 1 : ((p.C % (64 * V) == 0) ? 2 : 1);
    TileMap tm;
    int total = 0, ns = 0;
    tm.groups = (p.C + 32 * V * nv - 1) / (32 * V * nv);
    for (int l = p.L - 1; l >= 0; --l) {
      if (heavy(l) != (pass == 0)) continue;
      tm.start[ns] = total;
      tm.level[ns] = l;
      tm.tiles_x[ns] = (p.W[l] + tw - 1) / tw;
      tm.tiles_y[ns] = (p.H[l] + kTileH - 1) / kTileH;
      total += tm.tiles_x[ns] * tm.tiles_y[ns] * p.B * tm.groups;
      ++ns;
    }
    tm.nslots = ns;
    for (int j2 = ns; j2 <= kMaxLevels; ++j2) tm.start[j2] = total;
    for (int j2 = ns; j2 < kMaxLevels; ++j2) { tm.level[j2] = 0; tm.tiles_x[j2] = tm.tiles_y[j2] = 1; }
    if (total == 0) continue;
    if (pass == 0) {
      if (dtype == 0) roi_bwd_pull<float, 1, 4><<<total, kThreads, 0, stream>>>(p, ws, tm);
      else roi_bwd_pull<__nv_bfloat16, 1, 4><<<total, kThreads, 0, stream>>>(p, ws, tm);
    } else if (nv == 2) {
      roi_bwd_pull<float, 2, 8><<<total, kThreads, 0, stream>>>(p, ws, tm);
    } else {
      if (dtype == 0) roi_bwd_pull<float, 1, 8><<<total, kThreads, 0, stream>>>(p, ws, tm);
      else roi_bwd_pull<__nv_bfloat16, 1, 8><<<total, kThreads, 0, stream>>>(p, ws, tm);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Device pointer to the per-region headers inside a laid-out workspace (the
// atomic fallback kernel reads the `flags` field).
const void* roi_pull_headers(int K, int R, int L, int B, const int* H, void* workspace) {
  PullWs ws;
  pull_ws_layout(K * R, L, B, H, static_cast<unsigned char*>(workspace), &ws);
  return ws.hdr;
}

}  // namespace arfe
