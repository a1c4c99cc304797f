// AR-RFF extraction: region generation + RoI->level map + multi-level RoIAlign
// fused into one launch that writes the concatenated [K, R*C, PH, PW] tensor.
//
// Replaces (reference, relative to /root/reference):
//   mmdet/models/utils/additional.py:38-71            get_adaptive_scale_rois
//   mmdet/models/roi_heads/roi_extractors/single_level.py:53-152
//   mmdet/ops/roi_align/src/cuda/roi_align_kernel_v2.cu:62-128, :179-263
//   mmdet/models/roi_heads/standard_roi_head.py:138-155   (3x extract + cat)
//
// Design: one CTA per (RoI, region).  The bilinear
// sampling pattern of RoIAlign is separable, so the CTA first folds the
// gh x gw samples of every bin into two small per-axis tables in shared
// memory -- for bin row ph: first feature row, number of rows, and one
// aggregated weight per row (same for columns).  A bin is then
//     out[ph][pw] = (1/count) * sum_rows wy[row] * sum_cols wx[col] * f[row][col]
// which touches each feature value once per bin (~12 taps instead of the
// reference's 4 * gh * gw ~ 40 loads per output element) and is shared by all
// C channels.
//   NCHW: the RoI's feature window is staged through shared memory in
//         [pixel][channel] order (pitch 33 -> conflict-free both ways): global
//         reads are coalesced along x, compute runs lane == channel, and the
//         finished [channel][bin] block is written with coalesced 128-bit rows.
//   NHWC: lanes == channels read global memory directly with 128-bit loads.
// RoIs whose tables or window do not fit the fixed shared-memory budget take a
// slow generic path (direct taps, reference loop order).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "roi_common.cuh"

namespace arfe {

// ---------------------------------------------------------------------------
// Forward, NCHW features.
// dynamic smem: [CtaHeader][AxisTable y][AxisTable x][win: kCapPx*33][outs: 32*opitch]
//
// Per band (a run of bin rows whose feature rows fit the staged window, normally
// the whole RoI) every lane precomputes the global offsets of "its" <= 15 window
// pixels once; the staging loop over channels is then one address add + one
// cp.async (LDGSTS, no register round trip) per element, with up to 60 copies
// in flight per thread.  Compute runs lane == channel out of the [pixel][33]
// window with the bin's column weights held in registers and the column loop
// unrolled (template NC).
// ---------------------------------------------------------------------------
constexpr int kCapPx = 480;               // staged window pixels per band
constexpr int kStageIters = kCapPx / 32;  // pixels per lane

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Stage N*32 window pixels of one channel plane with N loads in flight, then N
// conflict-free stores ([pixel][33] layout).  4-byte cp.async (LDGSTS) was
// measured slower here: it occupies the LSU/MIO pipe ~8 cycles per warp-copy,
// LDG + STS about half of that.  No predicates: offsets past the window are
// clamped to a valid pixel and land in unused slots of the staging buffer.
template <typename T, int N>
__device__ __forceinline__ void stage_plane(const T* __restrict__ plane, float* __restrict__ dst,
                                            const unsigned (&goff)[kStageIters]) {
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = to_f(__ldg(plane + goff[i]));
#pragma unroll
  for (int i = 0; i < N; ++i) dst[i * 32 * kPitch] = v[i];
}

template <typename T>
__device__ __forceinline__ void stage_plane_n(int n, const T* __restrict__ plane,
                                              float* __restrict__ dst,
                                              const unsigned (&goff)[kStageIters]) {
  switch (n) {
#define ARFE_CASE(N) case N: stage_plane<T, N>(plane, dst, goff); break;
    ARFE_CASE(1) ARFE_CASE(2) ARFE_CASE(3) ARFE_CASE(4) ARFE_CASE(5) ARFE_CASE(6) ARFE_CASE(7)
    ARFE_CASE(8) ARFE_CASE(9) ARFE_CASE(10) ARFE_CASE(11) ARFE_CASE(12) ARFE_CASE(13) ARFE_CASE(14)
    ARFE_CASE(15)
#undef ARFE_CASE
    default: break;
  }
}

// Per-band descriptors (built once per band, read by every chunk).
struct BandDesc {
  int4 ph[kMaxPool];  // x: window offset of the bin row's first feature row (floats), y: rows, z: offset of wy
  int4 pw[kMaxPool];  // x: window offset of the bin column's first feature column, y: cols, z: offset of wx
};

// All bins of one bin column `pw` (weights in registers), rows ph0..ph1, for
// the 32 channels of the staged chunk (lane == channel).
template <int NC>
__device__ __forceinline__ void column_bins(const float* __restrict__ win, const BandDesc& bd,
                                            const AxisTable& ty, const float* __restrict__ wx,
                                            int coloff, int rowpitch, int nph, int lane,
                                            float inv_count, float* __restrict__ outs_lane,
                                            int PW, int pw) {
  float w[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) w[j] = wx[j];
  const float* __restrict__ colbase = win + coloff + lane;
  for (int q = 0; q < nph; ++q) {
    const int4 d = bd.ph[q];
    const float* __restrict__ p = colbase + d.x;
    const float* __restrict__ wy = ty.w + d.z;
    float acc = 0.f;
    for (int jr = 0; jr < d.y; ++jr) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < NC; ++j) t = fmaf(w[j], p[j * kPitch], t);
      acc = fmaf(wy[jr], t, acc);
      p += rowpitch;
    }
    outs_lane[q * PW + pw] = acc * inv_count;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
roi_fuse_fwd_nchw(const RoiFuseParams p, int opitch) {
  extern __shared__ __align__(16) unsigned char smem[];
  CtaHeader& hd = *reinterpret_cast<CtaHeader*>(smem);
  AxisTable& ty = *reinterpret_cast<AxisTable*>(smem + 128);
  AxisTable& tx = *reinterpret_cast<AxisTable*>(smem + 128 + sizeof(AxisTable));
  BandDesc& bd = *reinterpret_cast<BandDesc*>(smem + 128 + 2 * sizeof(AxisTable));
  float* win = reinterpret_cast<float*>(smem + 128 + 2 * sizeof(AxisTable) + sizeof(BandDesc));
  float* outs = win + (size_t)kCapPx * kPitch;

  const int k = blockIdx.x / p.R, r = blockIdx.x % p.R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthr = blockDim.x, nw = nthr >> 5;
  const int PH = p.PH, PW = p.PW, PHW = PH * PW, C = p.C;
  T* __restrict__ out_blk =
      static_cast<T*>(p.out) + ((size_t)k * p.R + r) * C * PHW;

  if (!setup_cta(p, k, r, hd, ty, tx)) {
    for (int i = tid; i < C * PHW; i += nthr) out_blk[i] = from_f<T>(0.f);
    return;
  }
  const int ww = hd.xmax - hd.xmin + 1;
  if (hd.overflow || (long long)hd.max_rows * ww > kCapPx) {
    forward_generic<T, false>(p, hd, out_blk);
    return;
  }
  if (ARFE_SKIP(p, 8)) return;  // setup only
  const int H = hd.H, W = hd.W;
  const size_t HW = (size_t)H * W;
  const float inv_count = 1.0f / hd.g.count;
  const T* __restrict__ fimg =
      static_cast<const T*>(p.feats[hd.lvl]) + (size_t)hd.g.batch * C * HW;
  const float inv_ww = 1.0f / (float)ww;
  const int rowpitch = ww * kPitch;

  int ph0 = 0;
  while (ph0 < PH) {
    // Greedy band [ph0, ph1): as many bin rows as fit the staged window.
    int r0 = 0x7fffffff, r1 = -1, ph1 = ph0;
    while (ph1 < PH) {
      int a = r0, b = r1;
      if (ty.cnt[ph1] > 0) {
        a = min(a, ty.first[ph1]);
        b = max(b, ty.first[ph1] + ty.cnt[ph1] - 1);
      }
      if (b >= a && (b - a + 1) * ww > kCapPx) break;
      r0 = a; r1 = b; ++ph1;
    }
    const int npx = (r1 >= r0) ? (r1 - r0 + 1) * ww : 0;
    const int nph = ph1 - ph0;
    const int run = nph * PW;  // contiguous output floats per channel for this band

    __syncthreads();  // previous band's readers of bd/outs are done
    if (tid < nph) {
      const int q = ph0 + tid;
      bd.ph[tid] = make_int4(ty.cnt[q] > 0 ? (ty.first[q] - r0) * rowpitch : 0, ty.cnt[q], ty.off[q], 0);
    } else if (tid >= 32 && tid < 32 + PW) {
      const int q = tid - 32;
      bd.pw[q] = make_int4(tx.cnt[q] > 0 ? (tx.first[q] - hd.xmin) * kPitch : 0, tx.cnt[q], tx.off[q], 0);
    }

    // global offsets (within a channel plane) of this lane's window pixels;
    // pixels past the window (last iteration only) re-read the first pixel
    unsigned goff[kStageIters];
    const int niter = (npx + 31) >> 5;
#pragma unroll
    for (int i = 0; i < kStageIters; ++i) {
      const int px = lane + 32 * i;
      int row = 0, col = 0;
      if (px < npx) split_px(px, ww, inv_ww, row, col);
      goff[i] = (unsigned)((r0 + row) * W + hd.xmin + col);
    }
    if (npx == 0) {
#pragma unroll
      for (int i = 0; i < kStageIters; ++i) goff[i] = 0;
    }

    for (int c0 = 0; c0 < C; c0 += kChunk) {
      const int cc = min(kChunk, C - c0);
      // ---- stage the band's window for channels c0..c0+cc ----
      for (int c = warp; c < cc; c += nw) {
        const T* __restrict__ plane = fimg + (size_t)(c0 + c) * HW;
        float* __restrict__ dst = win + lane * kPitch + c;
        if (!ARFE_SKIP(p, 2)) stage_plane_n<T>(niter, plane, dst, goff);
      }
      __syncthreads();
      // ---- compute: one warp per bin column, lane == channel ----
      if (lane < cc && !ARFE_SKIP(p, 1)) {
        float* __restrict__ outs_lane = opitch > 0 ? outs + lane * opitch : nullptr;
        for (int pw = warp; pw < PW; pw += nw) {
          const int4 d = bd.pw[pw];
          const float* __restrict__ wx = tx.w + d.z;
          if (opitch > 0 && d.y >= 1 && d.y <= 6) {
            switch (d.y) {
              case 1: column_bins<1>(win, bd, ty, wx, d.x, rowpitch, nph, lane, inv_count, outs_lane, PW, pw); break;
              case 2: column_bins<2>(win, bd, ty, wx, d.x, rowpitch, nph, lane, inv_count, outs_lane, PW, pw); break;
              case 3: column_bins<3>(win, bd, ty, wx, d.x, rowpitch, nph, lane, inv_count, outs_lane, PW, pw); break;
              case 4: column_bins<4>(win, bd, ty, wx, d.x, rowpitch, nph, lane, inv_count, outs_lane, PW, pw); break;
              case 5: column_bins<5>(win, bd, ty, wx, d.x, rowpitch, nph, lane, inv_count, outs_lane, PW, pw); break;
              default: column_bins<6>(win, bd, ty, wx, d.x, rowpitch, nph, lane, inv_count, outs_lane, PW, pw); break;
            }
          } else {
            for (int q = 0; q < nph; ++q) {
              const int4 e = bd.ph[q];
              const float* __restrict__ pp = win + d.x + e.x + lane;
              float acc = 0.f;
              for (int jr = 0; jr < e.y; ++jr) {
                float t = 0.f;
                for (int j = 0; j < d.y; ++j) t = fmaf(wx[j], pp[j * kPitch], t);
                acc = fmaf(ty.w[e.z + jr], t, acc);
                pp += rowpitch;
              }
              const float v = acc * inv_count;
              if (opitch > 0) outs_lane[q * PW + pw] = v;
              else out_blk[(size_t)(c0 + lane) * PHW + (ph0 + q) * PW + pw] = from_f<T>(v);
            }
          }
        }
      }
      __syncthreads();
      if (opitch > 0 && !ARFE_SKIP(p, 4)) {
        // ---- coalesced write: per channel a run of `run` floats at bin ph0*PW ----
        T* __restrict__ dst = out_blk + (size_t)c0 * PHW + ph0 * PW;
        for (int c = warp; c < cc; c += nw)
          for (int b = lane; b < run; b += 32)
            dst[(size_t)c * PHW + b] = from_f<T>(outs[c * opitch + b]);
        // the next chunk's post-staging __syncthreads orders the reuse of outs
      }
    }
    ph0 = ph1;
  }
}

// ---------------------------------------------------------------------------
// Forward, NHWC features: lanes == channels straight from global memory.
// dynamic smem: [CtaHeader][AxisTable y][AxisTable x][outs: C*opitch]
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
roi_fuse_fwd_nhwc(const RoiFuseParams p, int opitch) {
  extern __shared__ __align__(16) unsigned char smem[];
  CtaHeader& hd = *reinterpret_cast<CtaHeader*>(smem);
  AxisTable& ty = *reinterpret_cast<AxisTable*>(smem + 128);
  AxisTable& tx = *reinterpret_cast<AxisTable*>(smem + 128 + sizeof(AxisTable));
  float* outs = reinterpret_cast<float*>(smem + 128 + 2 * sizeof(AxisTable));

  const int k = blockIdx.x / p.R, r = blockIdx.x % p.R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int PH = p.PH, PW = p.PW, PHW = PH * PW, C = p.C;
  T* __restrict__ out_blk =
      static_cast<T*>(p.out) + ((size_t)k * p.R + r) * C * PHW;

  if (!setup_cta(p, k, r, hd, ty, tx)) {
    zero_block(out_blk, C * PHW);
    return;
  }
  if (hd.overflow || opitch <= 0) {
    forward_generic<T, true>(p, hd, out_blk);
    return;
  }
  const int H = hd.H, W = hd.W;
  const float count = hd.g.count;
  const T* __restrict__ fimg =
      static_cast<const T*>(p.feats[hd.lvl]) + (size_t)hd.g.batch * H * W * C;

  for (int bin = warp; bin < PHW; bin += kWarps) {
    const int ph = bin / PW, pw = bin % PW;
    const int nr = ty.cnt[ph], nc = tx.cnt[pw];
    const float* __restrict__ wy = ty.w + ty.off[ph];
    const float* __restrict__ wx = tx.w + tx.off[pw];
    const T* __restrict__ base =
        fimg + ((size_t)ty.first[ph] * W + tx.first[pw]) * C;
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int jr = 0; jr < nr; ++jr) {
        const T* __restrict__ prow = base + (size_t)jr * W * C + c;
        float t = 0.f;
        for (int jc = 0; jc < nc; ++jc)
          t = fmaf(wx[jc], to_f(prow[(size_t)jc * C]), t);
        acc = fmaf(wy[jr], t, acc);
      }
      outs[c * opitch + bin] = __fdiv_rn(acc, count);
    }
  }
  __syncthreads();
  const int total = C * PHW;
  for (int e = tid; e < total; e += kThreads) {
    const int c = e / PHW;
    const int b = e - c * PHW;
    out_blk[e] = from_f<T>(outs[c * opitch + b]);
  }
}


// Backward inner loops: one window pixel per lane, walk the channels.  PHW_T is
// the compile-time bins-per-channel (0 = runtime) so the four taps of every
// unrolled channel are LDS with immediate offsets; the destination pointer
// advances by the channel stride.
template <int PHW_T>
__device__ __forceinline__ void bwd_px_scalar(const float* __restrict__ dsm, int PHW, int cc,
                                              float w00, float w01, float w10, float w11,
                                              int i00, int i01, int i10, int i11,
                                              float* __restrict__ dst, size_t cstride) {
  const int phw = PHW_T ? PHW_T : PHW;
  const float* __restrict__ d0 = dsm + i00;
  const float* __restrict__ d1 = dsm + i01;
  const float* __restrict__ d2 = dsm + i10;
  const float* __restrict__ d3 = dsm + i11;
#pragma unroll 8
  for (int c = 0; c < cc; ++c) {
    const float v = w00 * d0[c * phw] + w01 * d1[c * phw] + w10 * d2[c * phw] + w11 * d3[c * phw];
    atomicAdd(dst, v);
    dst += cstride;
  }
}

// ---------------------------------------------------------------------------
// Backward.  Thread == one pixel of the RoI's feature window; it looks up the
// (usually <= 2 x 2) bins that sample it once, keeps their weights in
// registers, and then walks the channels: a few shared-memory reads of the
// staged dout block and ONE reduction into the gradient map per (pixel,
// channel) -- coalesced along x for NCHW -- instead of the reference's
// 4 * gh * gw atomics per output element.
// dynamic smem: [CtaHeader][AxisTable y][AxisTable x][dsm: cb*PHW floats]
// ---------------------------------------------------------------------------
template <typename T, bool kNHWC>
__device__ __forceinline__ void roi_fuse_bwd_region(const RoiFuseParams& p, int cb, int region,
                                                    unsigned char* smem) {
  CtaHeader& hd = *reinterpret_cast<CtaHeader*>(smem);
  AxisTable& ty = *reinterpret_cast<AxisTable*>(smem + 128);
  AxisTable& tx = *reinterpret_cast<AxisTable*>(smem + 128 + sizeof(AxisTable));
  float* dsm = reinterpret_cast<float*>(smem + 128 + 2 * sizeof(AxisTable));

  const int k = region / p.R, r = region % p.R;
  const int tid = threadIdx.x;
  const int PH = p.PH, PW = p.PW, PHW = PH * PW, C = p.C;
  const T* __restrict__ dout_blk =
      static_cast<const T*>(p.dout) + ((size_t)k * p.R + r) * C * PHW;
  // element (c, bin) of this region's incoming gradient, NCHW or channels-last
  auto dout_at = [&](int c, int bin) -> float {
    return p.dout_cl ? to_f(static_cast<const T*>(p.dout)[((size_t)k * PHW + bin) * p.bin_stride + p.reg_off[r] + c])
                     : to_f(dout_blk[(size_t)c * PHW + bin]);
  };

  if (!setup_cta(p, k, r, hd, ty, tx)) return;
  const int H = hd.H, W = hd.W;
  const RoiGeom g = hd.g;
  float* __restrict__ dimg = p.dfeats[hd.lvl] + (size_t)g.batch * C * H * W;

  if (hd.overflow) {
    // generic path: reference loop order, 4 atomics per sample
    for (int e = tid; e < C * PHW; e += kThreads) {
      const int c = e / PHW, bin = e - c * PHW;
      const int ph = bin / PW, pw = bin % PW;
      const float gv = dout_at(c, bin);
      for (int iy = 0; iy < g.grid_h; ++iy) {
        AxisTap a = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H);
        if (a.lo < 0) continue;
        for (int ix = 0; ix < g.grid_w; ++ix) {
          AxisTap b = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W);
          if (b.lo < 0) continue;
          const float s = gv / g.count;
          if (kNHWC) {
            atomicAdd(dimg + ((size_t)a.lo * W + b.lo) * C + c, s * a.wl * b.wl);
            atomicAdd(dimg + ((size_t)a.lo * W + b.hi) * C + c, s * a.wl * b.wh);
            atomicAdd(dimg + ((size_t)a.hi * W + b.lo) * C + c, s * a.wh * b.wl);
            atomicAdd(dimg + ((size_t)a.hi * W + b.hi) * C + c, s * a.wh * b.wh);
          } else {
            float* pl = dimg + (size_t)c * H * W;
            atomicAdd(pl + (size_t)a.lo * W + b.lo, s * a.wl * b.wl);
            atomicAdd(pl + (size_t)a.lo * W + b.hi, s * a.wl * b.wh);
            atomicAdd(pl + (size_t)a.hi * W + b.lo, s * a.wh * b.wl);
            atomicAdd(pl + (size_t)a.hi * W + b.hi, s * a.wh * b.wh);
          }
        }
      }
    }
    return;
  }

  // (128-bit vector reductions over x were measured: no gain -- the L2 atomic
  // units retire ~1 fp32 element per slice-clock whatever the request width.)
  const int xa = hd.xmin;
  const int ww = hd.xmax - hd.xmin + 1, wh = hd.ymax - hd.ymin + 1;
  const int npx = ww * wh;
  const float inv_ww = 1.0f / (float)ww;
  const float inv_count = 1.0f / g.count;
  const int lane = tid & 31;
  const size_t HW = (size_t)H * W;

  for (int c0 = 0; c0 < C; c0 += cb) {
    const int cc = min(cb, C - c0);
    __syncthreads();
    if (p.dout_cl) {
      for (int e = tid; e < cc * PHW; e += kThreads) {
        const int bin = e / cc, c = e - bin * cc;
        dsm[c * PHW + bin] = dout_at(c0 + c, bin);
      }
    } else {
      for (int e = tid; e < cc * PHW; e += kThreads)
        dsm[e] = to_f(dout_blk[(size_t)c0 * PHW + e]);
    }
    __syncthreads();

    int complex_px = 0;  // this thread saw a pixel sampled by more than 2 x 2 bins
    for (int px0 = tid - lane; px0 < npx; px0 += kThreads) {  // warp-uniform bound
      const int px = px0 + lane;
      int wr = 0, wc = 0;
      if (px < npx) split_px(px, ww, inv_ww, wr, wc);
      const int row = hd.ymin + wr, col = xa + wc;
      // bins covering this row / column (contiguous ranges)
      int pa = -1, na = 0, pb = -1, nb = 0;
      if (px < npx) {
        for (int q = 0; q < PH; ++q)
          if (ty.cnt[q] > 0 && row >= ty.first[q] && row < ty.first[q] + ty.cnt[q]) {
            if (pa < 0) pa = q;
            na = q - pa + 1;
          }
        for (int q = 0; q < PW; ++q)
          if (tx.cnt[q] > 0 && col >= tx.first[q] && col < tx.first[q] + tx.cnt[q]) {
            if (pb < 0) pb = q;
            nb = q - pb + 1;
          }
      }
      const bool live = na > 0 && nb > 0;
      float* __restrict__ dst =
          kNHWC ? dimg + ((size_t)row * W + col) * C + c0
                : dimg + (size_t)c0 * HW + (size_t)row * W + col;

      auto wy_of = [&](int q) -> float {
        const int j = row - ty.first[q];
        return (ty.cnt[q] > 0 && j >= 0 && j < ty.cnt[q]) ? ty.w[ty.off[q] + j] : 0.f;
      };
      auto wx_of = [&](int q) -> float {
        const int j = col - tx.first[q];
        return (tx.cnt[q] > 0 && j >= 0 && j < tx.cnt[q]) ? tx.w[tx.off[q] + j] : 0.f;
      };
      if (live && (na > 2 || nb > 2)) complex_px = 1;
      if (live && na <= 2 && nb <= 2) {
        const int pa1 = min(pa + 1, PH - 1), pb1 = min(pb + 1, PW - 1);
        const float wa0 = wy_of(pa) * inv_count;
        const float wa1 = (na > 1) ? wy_of(pa + 1) * inv_count : 0.f;
        const float wb0 = wx_of(pb);
        const float wb1 = (nb > 1) ? wx_of(pb + 1) : 0.f;
        const float w00 = wa0 * wb0, w01 = wa0 * wb1, w10 = wa1 * wb0, w11 = wa1 * wb1;
        const int i00 = pa * PW + pb, i01 = pa * PW + pb1, i10 = pa1 * PW + pb, i11 = pa1 * PW + pb1;
        if (kNHWC) {
          int c = 0;
          if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (; c + 3 < cc; c += 4) {
              float v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float* __restrict__ d = dsm + (c + u) * PHW;
                v[u] = w00 * d[i00] + w01 * d[i01] + w10 * d[i10] + w11 * d[i11];
              }
              red_add_v4(dst + c, v[0], v[1], v[2], v[3]);
            }
          }
          for (; c < cc; ++c) {
            const float* __restrict__ d = dsm + c * PHW;
            atomicAdd(dst + c, w00 * d[i00] + w01 * d[i01] + w10 * d[i10] + w11 * d[i11]);
          }
        } else if (PHW == 49) {
          bwd_px_scalar<49>(dsm, PHW, cc, w00, w01, w10, w11, i00, i01, i10, i11, dst, HW);
        } else {
          bwd_px_scalar<0>(dsm, PHW, cc, w00, w01, w10, w11, i00, i01, i10, i11, dst, HW);
        }
      }
      // rows / columns sampled by more than two bins are finished below
    }
    // ---- generic pass: pixels sampled by more than 2 x 2 bins, one (pixel,
    // channel) pair per thread so that tiny windows still use the whole CTA ----
    if (!__syncthreads_or(complex_px)) continue;
    for (int item = tid; item < npx * cc; item += kThreads) {
      const int c = item / npx, px = item - c * npx;
      int wr, wc;
      split_px(px, ww, inv_ww, wr, wc);
      const int row = hd.ymin + wr, col = xa + wc;
      int pa = -1, na = 0, pb = -1, nb = 0;
      for (int q = 0; q < PH; ++q)
        if (ty.cnt[q] > 0 && row >= ty.first[q] && row < ty.first[q] + ty.cnt[q]) {
          if (pa < 0) pa = q;
          na = q - pa + 1;
        }
      for (int q = 0; q < PW; ++q)
        if (tx.cnt[q] > 0 && col >= tx.first[q] && col < tx.first[q] + tx.cnt[q]) {
          if (pb < 0) pb = q;
          nb = q - pb + 1;
        }
      if (na == 0 || nb == 0 || (na <= 2 && nb <= 2)) continue;
      const float* __restrict__ d = dsm + c * PHW;
      float v = 0.f;
      for (int a = 0; a < na; ++a) {
        const int ja = row - ty.first[pa + a];
        const float wa = (ty.cnt[pa + a] > 0 && ja >= 0 && ja < ty.cnt[pa + a]) ? ty.w[ty.off[pa + a] + ja] : 0.f;
        float t = 0.f;
        for (int b = 0; b < nb; ++b) {
          const int jb = col - tx.first[pb + b];
          const float wb = (tx.cnt[pb + b] > 0 && jb >= 0 && jb < tx.cnt[pb + b]) ? tx.w[tx.off[pb + b] + jb] : 0.f;
          t = fmaf(wb, d[(pa + a) * PW + pb + b], t);
        }
        v = fmaf(wa, t, v);
      }
      float* __restrict__ dst = kNHWC ? dimg + ((size_t)row * W + col) * C + c0 + c
                                      : dimg + (size_t)(c0 + c) * HW + (size_t)row * W + col;
      atomicAdd(dst, v * inv_count);
    }
  }
}

// Grid = one CTA per (RoI, region); or, as the fallback behind the pull kernel
// (p.flag_list / p.flag_count: region ids written by roi_prep_kernel), a
// fixed small grid walking the regions whose tables did not fit.
template <typename T, bool kNHWC>
__global__ void __launch_bounds__(kThreads, 3)
roi_fuse_bwd(const RoiFuseParams p, int cb /* channels staged per pass */) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (p.flag_list) {
    const int n = *p.flag_count;
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
      roi_fuse_bwd_region<T, kNHWC>(p, cb, p.flag_list[j], smem);
      __syncthreads();
    }
  } else {
    roi_fuse_bwd_region<T, kNHWC>(p, cb, blockIdx.x, smem);
  }
}

// ---------------------------------------------------------------------------
// Parity instrumentation: dump boxes / levels / grids / taps.
// ---------------------------------------------------------------------------
__global__ void roi_fuse_taps_kernel(const RoiFuseParams p, int max_grid,
                                     int32_t* lvl, int32_t* grid, float* boxes,
                                     int32_t* ylo, int32_t* yhi, float* ywl,
                                     float* ywh, int32_t* xlo, int32_t* xhi,
                                     float* xwl, float* xwh) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.K * p.R) return;
  const int r = idx / p.K, k = idx % p.K;
  RegionBox bx = region_box(p.rois + 5 * (size_t)k, r, p.facs);
  const int l = (p.L == 1) ? 0 : map_roi_level(bx, p.L, p.finest_scale);
  if (lvl) lvl[idx] = l;
  if (boxes) {
    float* o = boxes + (size_t)idx * 5;
    o[0] = bx.b; o[1] = bx.x1; o[2] = bx.y1; o[3] = bx.x2; o[4] = bx.y2;
  }
  const int ll = l < 0 ? 0 : l;
  RoiGeom g = roi_geometry(bx, p.scale[ll], p.PH, p.PW, p.sampling_ratio);
  if (grid) { grid[2 * idx] = g.grid_h; grid[2 * idx + 1] = g.grid_w; }
  if (ylo)
    for (int q = 0; q < p.PH; ++q)
      for (int i = 0; i < max_grid; ++i) {
        const size_t o = ((size_t)idx * p.PH + q) * max_grid + i;
        if (i < g.grid_h) {
          AxisTap t = axis_sample(g.start_h, q, g.bin_h, i, g.grid_h, p.H[ll]);
          ylo[o] = t.lo; yhi[o] = t.hi; ywl[o] = t.wl; ywh[o] = t.wh;
        } else { ylo[o] = yhi[o] = -2; ywl[o] = ywh[o] = 0.f; }
      }
  if (xlo)
    for (int q = 0; q < p.PW; ++q)
      for (int i = 0; i < max_grid; ++i) {
        const size_t o = ((size_t)idx * p.PW + q) * max_grid + i;
        if (i < g.grid_w) {
          AxisTap t = axis_sample(g.start_w, q, g.bin_w, i, g.grid_w, p.W[ll]);
          xlo[o] = t.lo; xhi[o] = t.hi; xwl[o] = t.wl; xwh[o] = t.wh;
        } else { xlo[o] = xhi[o] = -2; xwl[o] = xwh[o] = 0.f; }
      }
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
constexpr int kFwdHdrBytes = kHdrBytes + (int)sizeof(BandDesc);

cudaError_t launch_roi_fuse_forward(const RoiFuseParams& p, int dtype, int layout,
                                    cudaStream_t stream) {
  const int PHW = p.PH * p.PW;
  const int grid = p.K * p.R;
  cudaError_t e;
  if (layout == 0) {
    int opitch = (PHW <= 256) ? (PHW | 1) : 0;  // staged output block or direct stores
    const int smem = kFwdHdrBytes + kCapPx * kPitch * 4 + kChunk * opitch * 4;
    const int threads = (p.PW % 7 == 0) ? 224 : kThreads;  // one warp per bin column
    if (dtype == 0) {
      if ((e = set_smem(roi_fuse_fwd_nchw<float>, smem)) != cudaSuccess) return e;
      roi_fuse_fwd_nchw<float><<<grid, threads, smem, stream>>>(p, opitch);
    } else {
      if ((e = set_smem(roi_fuse_fwd_nchw<__nv_bfloat16>, smem)) != cudaSuccess) return e;
      roi_fuse_fwd_nchw<__nv_bfloat16><<<grid, threads, smem, stream>>>(p, opitch);
    }
  } else if (p.C % (dtype == 0 ? 4 : 8) == 0) {
    return launch_roi_fuse_forward_cl(p, dtype, p.out_cl, stream);
  } else {
    if (p.out_cl) return cudaErrorInvalidValue;
    int opitch = PHW | 1;
    long long need = (long long)kHdrBytes + (long long)p.C * opitch * 4;
    if (need > kMaxSmem) { opitch = 0; need = kHdrBytes; }
    const int smem = (int)need;
    if (dtype == 0) {
      if ((e = set_smem(roi_fuse_fwd_nhwc<float>, smem)) != cudaSuccess) return e;
      roi_fuse_fwd_nhwc<float><<<grid, kThreads, smem, stream>>>(p, opitch);
    } else {
      if ((e = set_smem(roi_fuse_fwd_nhwc<__nv_bfloat16>, smem)) != cudaSuccess) return e;
      roi_fuse_fwd_nhwc<__nv_bfloat16><<<grid, kThreads, smem, stream>>>(p, opitch);
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_roi_fuse_backward(const RoiFuseParams& p, int dtype, int layout,
                                     cudaStream_t stream) {
  const int PHW = p.PH * p.PW;
  int grid = p.K * p.R;
  if (p.flag_list && grid > 296) grid = 296;  // fallback mode: usually nothing to do
  // stage as many channels of dout as fit ~100 KB (2 CTAs / SM)
  int cb = (100 * 1024 - kHdrBytes) / (PHW * 4);
  if (cb > p.C) cb = p.C;
  if (cb < 1) cb = 1;
  const int smem = kHdrBytes + cb * PHW * 4;
  if (smem > kMaxSmem) return cudaErrorInvalidValue;
  cudaError_t e;
#define ARFE_LAUNCH_BWD(TT, NH)                                              \
  do {                                                                       \
    if ((e = set_smem(roi_fuse_bwd<TT, NH>, smem)) != cudaSuccess) return e; \
    roi_fuse_bwd<TT, NH><<<grid, kThreads, smem, stream>>>(p, cb);           \
  } while (0)
  if (dtype == 0) {
    if (layout == 0) ARFE_LAUNCH_BWD(float, false); else ARFE_LAUNCH_BWD(float, true);
  } else {
    if (layout == 0) ARFE_LAUNCH_BWD(__nv_bfloat16, false); else ARFE_LAUNCH_BWD(__nv_bfloat16, true);
  }
#undef ARFE_LAUNCH_BWD
  return cudaGetLastError();
}

cudaError_t launch_roi_fuse_taps(const RoiFuseParams& p, int max_grid, int32_t* lvl,
                                 int32_t* grid, float* boxes, int32_t* ylo,
                                 int32_t* yhi, float* ywl, float* ywh, int32_t* xlo,
                                 int32_t* xhi, float* xwl, float* xwh,
                                 cudaStream_t stream) {
  const int n = p.K * p.R;
  roi_fuse_taps_kernel<<<(n + 127) / 128, 128, 0, stream>>>(
      p, max_grid, lvl, grid, boxes, ylo, yhi, ywl, ywh, xlo, xhi, xwl, xwh);
  return cudaGetLastError();
}

}  // namespace arfe
