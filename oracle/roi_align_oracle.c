/*
 * oracle/roi_align_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's CPU RoIAlign (aligned, "v2") used as
 * the checker for the CUDA kernels in arfe_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 *
 * What it restates (paths relative to /root/reference):
 *   forward   mmdet/ops/roi_align/src/cpu/roi_align_v2.cpp:107-191
 *             (+ the per-RoI tap pre-computation  :20-105)
 *   backward  mmdet/ops/roi_align/src/cpu/roi_align_v2.cpp:247-332
 *             (+ bilinear_interpolate_gradient    :193-240)
 *
 * Pinning: tests/test_oracle.py checks this file bit-for-bit against the
 * reference's own sources compiled unmodified (oracle/_ref, built by
 * oracle/build_oracle.py) and against torchvision.ops.roi_align, and against
 * the committed vectors in tests/golden/ that were produced by oracle/_ref.
 *
 * The float arithmetic below keeps the reference's operation ORDER (each
 * product/sum is a separate fp32 operation; x86-64 has no implicit FMA and
 * this file is compiled with -ffp-contract=off), because the integer results
 * derived from it -- grid sizes, tap rows/columns, validity -- must be
 * bit-exact.  The code structure is ours: taps are computed per axis
 * (separably) and combined, which yields identical numbers because a sample's
 * validity, indices and weights factor into a y-part and an x-part.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* One sample position along one axis (the reference computes the same
 * quantities inline, roi_align_v2.cpp:30-86 / :199-230). */
typedef struct {
  int lo, hi;   /* neighbouring rows (or columns); -1/-1 when outside */
  float wl, wh; /* weight of lo ("h*" in the reference) and of hi ("l*") */
} axis_tap;

static axis_tap axis_sample(float start, int p, float bin, int i, int grid,
                            int extent) {
  axis_tap t;
  /* roi_align_v2.cpp:30-32: start + p*bin + (i+.5f)*bin/grid, left to right */
  float pos = start + p * bin + (float)(i + .5f) * bin / (float)grid;
  if (pos < -1.0 || pos > extent) { /* :41 */
    t.lo = t.hi = -1;
    t.wl = t.wh = 0.f;
    return t;
  }
  if (pos <= 0) pos = 0; /* :57-62 */
  int lo = (int)pos;     /* :64 */
  int hi;
  if (lo >= extent - 1) { /* :69-74 */
    hi = lo = extent - 1;
    pos = (float)lo;
  } else {
    hi = lo + 1;
  }
  float l = pos - lo;            /* :83 */
  float h = (float)(1. - l);     /* :85 (double 1. - float, rounded to T) */
  t.lo = lo; t.hi = hi; t.wl = h; t.wh = l;
  return t;
}

typedef struct {
  int batch;
  float start_w, start_h, bin_w, bin_h;
  int grid_h, grid_w;
  int ok; /* 0 when the reference would raise (negative extent, :132-134) */
} roi_geom;

static roi_geom roi_geometry(const float* roi, float scale, int ph, int pw,
                             int sampling_ratio) {
  roi_geom g;
  g.batch = (int)roi[0];           /* :121 */
  const float off = 0.5f;          /* aligned=True, :124 */
  g.start_w = roi[1] * scale - off; /* :125-128 */
  g.start_h = roi[2] * scale - off;
  float end_w = roi[3] * scale - off;
  float end_h = roi[4] * scale - off;
  float rw = end_w - g.start_w;    /* :130-131 */
  float rh = end_h - g.start_h;
  g.ok = (rw >= 0 && rh >= 0);
  g.bin_h = rh / (float)ph;        /* :139-140 */
  g.bin_w = rw / (float)pw;
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceil(rh / ph); /* :143-147 */
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceil(rw / pw);
  return g;
}

/* Forward.  feat [B,C,H,W], rois [K,5], out [K,C,PH,PW]; returns 0, or -1 if
 * some RoI has negative extent (the reference asserts there). */
int oracle_roi_align_forward(const float* feat, int B, int C, int H, int W,
                             const float* rois, int K, float scale, int PH,
                             int PW, int sampling_ratio, float* out) {
  (void)B;
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 4)
  for (int n = 0; n < K; n++) {
    roi_geom g = roi_geometry(rois + 5 * n, scale, PH, PW, sampling_ratio);
    if (!g.ok) { rc = -1; continue; }
    int gh = g.grid_h > 0 ? g.grid_h : 0, gw = g.grid_w > 0 ? g.grid_w : 0;
    int prod = g.grid_h * g.grid_w;
    const float count = (float)(prod > 1 ? prod : 1); /* :151 */
    axis_tap* ty = (axis_tap*)malloc(sizeof(axis_tap) * (size_t)(PH * gh + 1));
    axis_tap* tx = (axis_tap*)malloc(sizeof(axis_tap) * (size_t)(PW * gw + 1));
    for (int p = 0; p < PH; p++)
      for (int i = 0; i < gh; i++)
        ty[p * gh + i] = axis_sample(g.start_h, p, g.bin_h, i, g.grid_h, H);
    for (int p = 0; p < PW; p++)
      for (int i = 0; i < gw; i++)
        tx[p * gw + i] = axis_sample(g.start_w, p, g.bin_w, i, g.grid_w, W);
    for (int c = 0; c < C; c++) {
      const float* plane = feat + ((size_t)g.batch * C + c) * H * W; /* :164-165 */
      float* o = out + ((size_t)n * C + c) * PH * PW;
      for (int p = 0; p < PH; p++)
        for (int q = 0; q < PW; q++) {
          float acc = 0.f; /* :172 */
          for (int iy = 0; iy < gh; iy++) {
            axis_tap a = ty[p * gh + iy];
            for (int ix = 0; ix < gw; ix++) {
              axis_tap b = tx[q * gw + ix];
              if (a.lo < 0 || b.lo < 0) {
                /* out-of-map sample: four zero weights at position 0 (:43-54);
                 * adds +0 terms, kept so NaN/Inf at plane[0] propagate alike */
                acc += 0.f * plane[0] + 0.f * plane[0] + 0.f * plane[0] +
                       0.f * plane[0];
                continue;
              }
              float w1 = a.wl * b.wl, w2 = a.wl * b.wh; /* :86 */
              float w3 = a.wh * b.wl, w4 = a.wh * b.wh;
              acc += w1 * plane[a.lo * W + b.lo] + w2 * plane[a.lo * W + b.hi] +
                     w3 * plane[a.hi * W + b.lo] + w4 * plane[a.hi * W + b.hi]; /* :176-179 */
            }
          }
          o[p * PW + q] = acc / count; /* :184 */
        }
    }
    free(ty);
    free(tx);
  }
  return rc;
}

/* Backward.  dout [K,C,PH,PW] contiguous, din [B,C,H,W] must be zero-filled by
 * the caller (the reference allocates at::zeros, :383-384).  Serial over
 * outputs in index order like the reference (:257), so sums are reproducible. */
int oracle_roi_align_backward(const float* dout, const float* rois, int K,
                              float scale, int PH, int PW, int B, int C, int H,
                              int W, int sampling_ratio, float* din) {
  (void)B;
  for (int n = 0; n < K; n++) {
    roi_geom g = roi_geometry(rois + 5 * n, scale, PH, PW, sampling_ratio);
    if (!g.ok) return -1;
    const float count = (float)(g.grid_h * g.grid_w); /* :300 (no max) */
    for (int c = 0; c < C; c++) {
      float* plane = din + ((size_t)g.batch * C + c) * H * W;
      const float* go = dout + ((size_t)n * C + c) * PH * PW;
      for (int p = 0; p < PH; p++)
        for (int q = 0; q < PW; q++) {
          const float gbin = go[p * PW + q];
          for (int iy = 0; iy < g.grid_h; iy++) {
            axis_tap a = axis_sample(g.start_h, p, g.bin_h, iy, g.grid_h, H);
            for (int ix = 0; ix < g.grid_w; ix++) {
              axis_tap b = axis_sample(g.start_w, q, g.bin_w, ix, g.grid_w, W);
              if (a.lo < 0 || b.lo < 0) continue; /* :318 guard, weights are 0 */
              float w1 = a.wl * b.wl, w2 = a.wl * b.wh;
              float w3 = a.wh * b.wl, w4 = a.wh * b.wh;
              plane[a.lo * W + b.lo] += gbin * w1 / count; /* :313-323 */
              plane[a.lo * W + b.hi] += gbin * w2 / count;
              plane[a.hi * W + b.lo] += gbin * w3 / count;
              plane[a.hi * W + b.hi] += gbin * w4 / count;
            }
          }
        }
    }
  }
  return 0;
}

/* Sampling-index dump used by the "indices bit-exact" parity test.
 * For every RoI: grid[n] = {grid_h, grid_w}; for every (p, i) with
 * i < min(grid, max_grid): ylo/yhi [K,PH,max_grid], xlo/xhi [K,PW,max_grid]
 * (-1 = sample outside the map, -2 = slot unused), plus the fp32 weights. */
int oracle_roi_align_taps(const float* rois, int K, float scale, int PH, int PW,
                          int H, int W, int sampling_ratio, int max_grid,
                          int32_t* grid, int32_t* ylo, int32_t* yhi,
                          float* ywl, float* ywh, int32_t* xlo, int32_t* xhi,
                          float* xwl, float* xwh) {
  for (int n = 0; n < K; n++) {
    roi_geom g = roi_geometry(rois + 5 * n, scale, PH, PW, sampling_ratio);
    grid[2 * n] = g.grid_h;
    grid[2 * n + 1] = g.grid_w;
    for (int p = 0; p < PH; p++)
      for (int i = 0; i < max_grid; i++) {
        size_t o = ((size_t)n * PH + p) * max_grid + i;
        if (i < g.grid_h) {
          axis_tap t = axis_sample(g.start_h, p, g.bin_h, i, g.grid_h, H);
          ylo[o] = t.lo; yhi[o] = t.hi; ywl[o] = t.wl; ywh[o] = t.wh;
        } else { ylo[o] = yhi[o] = -2; ywl[o] = ywh[o] = 0.f; }
      }
    for (int p = 0; p < PW; p++)
      for (int i = 0; i < max_grid; i++) {
        size_t o = ((size_t)n * PW + p) * max_grid + i;
        if (i < g.grid_w) {
          axis_tap t = axis_sample(g.start_w, p, g.bin_w, i, g.grid_w, W);
          xlo[o] = t.lo; xhi[o] = t.hi; xwl[o] = t.wl; xwh[o] = t.wh;
        } else { xlo[o] = xhi[o] = -2; xwl[o] = xwh[o] = 0.f; }
      }
  }
  return 0;
}
