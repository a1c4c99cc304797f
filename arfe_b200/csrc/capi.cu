// extern "C" surface of libarfe_b200.so (declared in include/arfe_b200.h):
// argument validation, error strings, parameter packing.  No allocation, no
// synchronisation, no global mutable state except the thread-local message.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/arfe_b200.h"
#include "launch.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_result(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return ARFE_OK;
  snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
size_t esize(int dtype) { return dtype == ARFE_F32 ? 4 : 2; }

// Launches go to the device that owns the tensors, not to whatever device happens to be
// current in the calling thread (the stream handle must belong to the same device).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* ptr) {
    cudaPointerAttributes a;
    if (ptr && cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
      if (cudaGetDevice(&prev) == cudaSuccess && a.device != prev)
        switched = cudaSetDevice(a.device) == cudaSuccess;
    } else {
      cudaGetLastError();  // not a device pointer: leave the error state clean, validation reports it
    }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define REQUIRE(cond, code, ...) \
  do { if (!(cond)) return fail(code, __VA_ARGS__); } while (0)

int check_common(const char* fn, int L, int B, int C, const int32_t* H, const int32_t* W,
                 int dtype, int layout) {
  REQUIRE(L >= 1 && L <= ARFE_MAX_LEVELS, ARFE_E_SHAPE, "%s: L=%d outside [1,%d]", fn, L, ARFE_MAX_LEVELS);
  REQUIRE(B >= 0 && C >= 1, ARFE_E_SHAPE, "%s: bad B=%d / C=%d", fn, B, C);
  REQUIRE(H && W, ARFE_E_NULL, "%s: H/W is NULL", fn);
  for (int l = 0; l < L; ++l)
    REQUIRE(H[l] >= 1 && W[l] >= 1 && H[l] < 32768 && W[l] < 32768, ARFE_E_SHAPE,
            "%s: level %d has size %dx%d (supported: 1..32767)", fn, l, H[l], W[l]);
  REQUIRE(dtype == ARFE_F32 || dtype == ARFE_BF16, ARFE_E_ENUM, "%s: unknown dtype %d", fn, dtype);
  REQUIRE(layout == ARFE_NCHW || layout == ARFE_NHWC, ARFE_E_ENUM, "%s: unknown layout %d", fn, layout);
  return ARFE_OK;
}

int fill_roi_params(const char* fn, arfe::RoiFuseParams& p, const int32_t* H, const int32_t* W,
                    const float* spatial_scale, int L, int B, int C, const float* rois, int K,
                    int regions, float facs, int PH, int PW, int sampling_ratio,
                    float finest_scale, int dtype, int layout) {
  int rc = check_common(fn, L, B, C, H, W, dtype, layout);
  if (rc) return rc;
  REQUIRE(spatial_scale, ARFE_E_NULL, "%s: spatial_scale is NULL", fn);
  REQUIRE(K >= 0, ARFE_E_SHAPE, "%s: K=%d", fn, K);
  REQUIRE(K == 0 || rois, ARFE_E_NULL, "%s: rois is NULL", fn);
  REQUIRE(regions == 1 || regions == 3, ARFE_E_ENUM, "%s: regions must be 1 or 3, got %d", fn, regions);
  REQUIRE(PH >= 1 && PH <= ARFE_MAX_POOL && PW >= 1 && PW <= ARFE_MAX_POOL, ARFE_E_SHAPE,
          "%s: pooled size %dx%d outside [1,%d]", fn, PH, PW, ARFE_MAX_POOL);
  REQUIRE(sampling_ratio >= 0, ARFE_E_SHAPE, "%s: sampling_ratio=%d", fn, sampling_ratio);
  REQUIRE(finest_scale > 0.f, ARFE_E_SHAPE, "%s: finest_scale must be > 0", fn);
  REQUIRE((long long)K * regions < (1ll << 31), ARFE_E_SHAPE, "%s: K*regions too large", fn);
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < L; ++l) { p.H[l] = H[l]; p.W[l] = W[l]; p.scale[l] = spatial_scale[l]; }
  p.L = L; p.B = B; p.C = C; p.K = K; p.R = regions; p.PH = PH; p.PW = PW;
  p.sampling_ratio = sampling_ratio; p.facs = facs; p.finest_scale = finest_scale;
  p.rois = rois;
  // channels-last RoI tensors: the concatenated layout [K][PH*PW][regions*C]
  for (int r = 0; r < 3; ++r) p.reg_off[r] = (long long)r * C;
  p.bin_stride = regions * C;
  return ARFE_OK;
}

// Split layout: one tensor [K][PH*PW][C] per region, anywhere in memory.
template <typename P>
int set_split_regions(const char* fn, arfe::RoiFuseParams& p, P const* reg, int regions,
                      int C, int dtype, const void** base) {
  REQUIRE(reg, ARFE_E_NULL, "%s: region pointer array is NULL", fn);
  for (int r = 0; r < regions; ++r) {
    REQUIRE(reg[r], ARFE_E_NULL, "%s: region %d is NULL", fn, r);
    REQUIRE(aligned(reg[r], 16), ARFE_E_ALIGN, "%s: region %d must be 16-byte aligned", fn, r);
    const long long d = static_cast<const char*>(reg[r]) - static_cast<const char*>(reg[0]);
    p.reg_off[r] = d / (long long)esize(dtype);
  }
  p.bin_stride = C;
  *base = reg[0];
  return ARFE_OK;
}

}  // namespace

extern "C" {

int arfe_version(void) { return ARFE_VERSION; }
const char* arfe_last_error(void) { return g_err; }

int arfe_roi_fuse_forward(const void* const* feats, const int32_t* H, const int32_t* W,
                          const float* spatial_scale, int L, int B, int C, const float* rois,
                          int K, int regions, float facs, int PH, int PW, int sampling_ratio,
                          float finest_scale, int dtype, int layout, int out_layout, void* out,
                          int32_t* lvl_out, float* boxes_out, void* stream) {
  const char* fn = "arfe_roi_fuse_forward";
  arfe::RoiFuseParams p;
  int rc = fill_roi_params(fn, p, H, W, spatial_scale, L, B, C, rois, K, regions, facs, PH, PW,
                           sampling_ratio, finest_scale, dtype, layout);
  if (rc) return rc;
  if (K == 0) return ARFE_OK;
  REQUIRE(feats && out, ARFE_E_NULL, "%s: feats/out is NULL", fn);
  REQUIRE(B >= 1, ARFE_E_SHAPE, "%s: B=0 with K>0", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(feats[l], ARFE_E_NULL, "%s: feats[%d] is NULL", fn, l);
    REQUIRE(aligned(feats[l], esize(dtype)), ARFE_E_ALIGN, "%s: feats[%d] misaligned", fn, l);
    p.feats[l] = feats[l];
  }
  REQUIRE(aligned(out, esize(dtype)) && aligned(rois, 4), ARFE_E_ALIGN, "%s: out/rois misaligned", fn);
  REQUIRE(out_layout == ARFE_NCHW || out_layout == ARFE_NHWC, ARFE_E_ENUM, "%s: unknown out_layout %d", fn, out_layout);
  if (out_layout == ARFE_NHWC) {
    REQUIRE(layout == ARFE_NHWC && C % (dtype == ARFE_F32 ? 4 : 8) == 0, ARFE_E_UNSUPPORTED,
            "%s: channels-last output needs channels-last features and C %% %d == 0", fn, dtype == ARFE_F32 ? 4 : 8);
  }
  if (layout == ARFE_NHWC && C % (dtype == ARFE_F32 ? 4 : 8) == 0) {
    for (int l = 0; l < L; ++l)
      REQUIRE(aligned(feats[l], 16), ARFE_E_ALIGN, "%s: feats[%d] must be 16-byte aligned", fn, l);
    REQUIRE(aligned(out, 16), ARFE_E_ALIGN, "%s: out must be 16-byte aligned", fn);
  }
  p.out = out; p.lvl_out = lvl_out; p.boxes_out = boxes_out; p.out_cl = out_layout == ARFE_NHWC;
  DeviceGuard guard(out);
  p.debug_skip = ARFE_KNOB_ENV("ARFE_FWD_SKIP", 0);
  return cuda_result(arfe::launch_roi_fuse_forward(p, dtype, layout, (cudaStream_t)stream), fn);
}

static int forward_plan_impl(const char* fn, const void* const* feats, const int32_t* H, const int32_t* W,
                             const float* spatial_scale, int L, int B, int C, const float* rois,
                             int K, int regions, float facs, int PH, int PW, int sampling_ratio,
                             float finest_scale, int dtype, void* out, void* const* out_regions,
                             int32_t* lvl_out, float* boxes_out, void* workspace,
                             size_t workspace_bytes, int stages, void* stream) {
  arfe::RoiFuseParams p;
  int rc = fill_roi_params(fn, p, H, W, spatial_scale, L, B, C, rois, K, regions, facs, PH, PW,
                           sampling_ratio, finest_scale, dtype, ARFE_NHWC);
  if (rc) return rc;
  if (K == 0) return ARFE_OK;
  REQUIRE(workspace, ARFE_E_NULL, "%s: workspace is NULL", fn);
  DeviceGuard guard(workspace);
  if (stages == 1) {  // plan only: no tensors involved
    REQUIRE(B >= 1 && aligned(rois, 4) && aligned(workspace, 256), ARFE_E_ALIGN, "%s: rois / workspace misaligned", fn);
    REQUIRE(workspace_bytes >= arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W), ARFE_E_SHAPE,
            "%s: workspace too small", fn);
    return cuda_result(arfe::launch_roi_fuse_forward_plan(p, dtype, workspace, workspace_bytes, 1, (cudaStream_t)stream), fn);
  }
  REQUIRE(feats && (out || out_regions), ARFE_E_NULL, "%s: feats/out is NULL", fn);
  REQUIRE(B >= 1, ARFE_E_SHAPE, "%s: B=0 with K>0", fn);
  REQUIRE(C % (dtype == ARFE_F32 ? 4 : 8) == 0, ARFE_E_UNSUPPORTED, "%s: C must be a multiple of %d",
          fn, dtype == ARFE_F32 ? 4 : 8);
  for (int l = 0; l < L; ++l) {
    REQUIRE(feats[l], ARFE_E_NULL, "%s: feats[%d] is NULL", fn, l);
    REQUIRE(aligned(feats[l], 16), ARFE_E_ALIGN, "%s: feats[%d] must be 16-byte aligned", fn, l);
    p.feats[l] = feats[l];
  }
  if (out_regions) {
    const void* base = nullptr;
    rc = set_split_regions(fn, p, out_regions, regions, C, dtype, &base);
    if (rc) return rc;
    out = const_cast<void*>(base);
  }
  REQUIRE(aligned(out, 16) && aligned(rois, 4) && aligned(workspace, 256), ARFE_E_ALIGN,
          "%s: out (16) / rois (4) / workspace (256) misaligned", fn);
  REQUIRE(workspace_bytes >= arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W), ARFE_E_SHAPE,
          "%s: workspace too small (%zu < %zu)", fn, workspace_bytes, arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W));
  p.out = out; p.lvl_out = lvl_out; p.boxes_out = boxes_out; p.out_cl = 1;
  p.debug_skip = ARFE_KNOB_ENV("ARFE_FWD_SKIP", 0);
  return cuda_result(arfe::launch_roi_fuse_forward_plan(p, dtype, workspace, workspace_bytes, stages, (cudaStream_t)stream), fn);
}

int arfe_roi_plan_build(const int32_t* H, const int32_t* W, const float* spatial_scale, int L, int B, int C,
                        const float* rois, int K, int regions, float facs, int PH, int PW,
                        int sampling_ratio, float finest_scale, int dtype, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return forward_plan_impl("arfe_roi_plan_build", nullptr, H, W, spatial_scale, L, B, C, rois, K, regions,
                           facs, PH, PW, sampling_ratio, finest_scale, dtype, nullptr, nullptr, nullptr,
                           nullptr, workspace, workspace_bytes, 1, stream);
}

int arfe_roi_fuse_forward_plan(const void* const* feats, const int32_t* H, const int32_t* W,
                               const float* spatial_scale, int L, int B, int C, const float* rois,
                               int K, int regions, float facs, int PH, int PW, int sampling_ratio,
                               float finest_scale, int dtype, void* out, int32_t* lvl_out,
                               float* boxes_out, void* workspace, size_t workspace_bytes,
                               int plan_ready, void* stream) {
  return forward_plan_impl("arfe_roi_fuse_forward_plan", feats, H, W, spatial_scale, L, B, C, rois, K,
                           regions, facs, PH, PW, sampling_ratio, finest_scale, dtype, out, nullptr,
                           lvl_out, boxes_out, workspace, workspace_bytes, plan_ready ? 2 : 3, stream);
}

int arfe_roi_fuse_forward_plan_split(const void* const* feats, const int32_t* H, const int32_t* W,
                                     const float* spatial_scale, int L, int B, int C,
                                     const float* rois, int K, int regions, float facs, int PH,
                                     int PW, int sampling_ratio, float finest_scale, int dtype,
                                     void* const* out_regions, void* workspace,
                                     size_t workspace_bytes, int plan_ready, void* stream) {
  if (!out_regions) return fail(ARFE_E_NULL, "arfe_roi_fuse_forward_plan_split: out_regions is NULL");
  return forward_plan_impl("arfe_roi_fuse_forward_plan_split", feats, H, W, spatial_scale, L, B, C, rois,
                           K, regions, facs, PH, PW, sampling_ratio, finest_scale, dtype, nullptr,
                           out_regions, nullptr, nullptr, workspace, workspace_bytes, plan_ready ? 2 : 3, stream);
}

int arfe_roi_fuse_backward(const void* dout, int dout_layout, const int32_t* H, const int32_t* W,
                           const float* spatial_scale, int L, int B, int C, const float* rois,
                           int K, int regions, float facs, int PH, int PW, int sampling_ratio,
                           float finest_scale, int dtype, int layout, float* const* dfeats,
                           void* stream) {
  const char* fn = "arfe_roi_fuse_backward";
  arfe::RoiFuseParams p;
  int rc = fill_roi_params(fn, p, H, W, spatial_scale, L, B, C, rois, K, regions, facs, PH, PW,
                           sampling_ratio, finest_scale, dtype, layout);
  if (rc) return rc;
  if (K == 0) return ARFE_OK;
  REQUIRE(dout && dfeats, ARFE_E_NULL, "%s: dout/dfeats is NULL", fn);
  REQUIRE(B >= 1, ARFE_E_SHAPE, "%s: B=0 with K>0", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(dfeats[l], ARFE_E_NULL, "%s: dfeats[%d] is NULL", fn, l);
    REQUIRE(aligned(dfeats[l], 4), ARFE_E_ALIGN, "%s: dfeats[%d] misaligned", fn, l);
    p.dfeats[l] = dfeats[l];
  }
  REQUIRE(dout_layout == ARFE_NCHW || dout_layout == ARFE_NHWC, ARFE_E_ENUM, "%s: unknown dout_layout %d", fn, dout_layout);
  p.dout = dout; p.dout_cl = dout_layout == ARFE_NHWC;
  DeviceGuard guard(dout);
  return cuda_result(arfe::launch_roi_fuse_backward(p, dtype, layout, (cudaStream_t)stream), fn);
}

size_t arfe_roi_plan_bytes(int K, int regions, int L, int B, const int32_t* H, const int32_t* W) {
  return arfe_roi_fuse_pull_workspace_bytes(K, regions, L, B, H, W);
}

size_t arfe_roi_fuse_pull_workspace_bytes(int K, int regions, int L, int B, const int32_t* H,
                                          const int32_t* W) {
  if (K <= 0 || (regions != 1 && regions != 3) || L < 1 || L > ARFE_MAX_LEVELS || B < 1 || !H || !W) return 0;
  for (int l = 0; l < L; ++l)
    if (H[l] < 1 || W[l] < 1) return 0;
  return arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W);
}

constexpr int kBinOnly = -77, kBinOnlySplit = -78;  // internal: arfe_roi_pull_bin (concatenated / split dout)

static int backward_pull_impl(const char* fn, const void* dout, const void* const* dout_regions,
                              const int32_t* H, const int32_t* W,
                                const float* spatial_scale, int L, int B, int C, const float* rois,
                                int K, int regions, float facs, int PH, int PW, int sampling_ratio,
                                float finest_scale, int dtype, float* const* dfeats,
                                void* workspace, size_t workspace_bytes, int plan_ready, void* stream) {
  const bool bin_split = plan_ready == kBinOnlySplit;
  if (bin_split) plan_ready = kBinOnly;
  arfe::RoiFuseParams p;
  int rc = fill_roi_params(fn, p, H, W, spatial_scale, L, B, C, rois, K, regions, facs, PH, PW,
                           sampling_ratio, finest_scale, dtype, ARFE_NHWC);
  if (rc) return rc;
  REQUIRE(dfeats || plan_ready == kBinOnly, ARFE_E_NULL, "%s: dfeats is NULL", fn);
  REQUIRE(B >= 1, ARFE_E_SHAPE, "%s: B=0", fn);
  REQUIRE(C % (dtype == ARFE_F32 ? 4 : 8) == 0, ARFE_E_UNSUPPORTED, "%s: C must be a multiple of %d",
          fn, dtype == ARFE_F32 ? 4 : 8);
  if (plan_ready == kBinOnly && K == 0) return ARFE_OK;
  DeviceGuard guard(workspace ? workspace : (dfeats ? static_cast<const void*>(dfeats[0]) : nullptr));
  for (int l = 0; l < L && dfeats; ++l) {
    REQUIRE(dfeats[l], ARFE_E_NULL, "%s: dfeats[%d] is NULL", fn, l);
    REQUIRE(aligned(dfeats[l], 16), ARFE_E_ALIGN, "%s: dfeats[%d] must be 16-byte aligned", fn, l);
    p.dfeats[l] = dfeats[l];
  }
  if (K == 0) {  // nothing is pooled: the gradient is all zeros, still fully written
    for (int l = 0; l < L; ++l) {
      cudaError_t e = cudaMemsetAsync(dfeats[l], 0, (size_t)B * C * H[l] * W[l] * 4, (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_result(e, fn);
    }
    return ARFE_OK;
  }
  const int stages = plan_ready == kBinOnly ? 2 : (plan_ready == 0 ? 7 : (plan_ready == 1 ? 6 : 4));
  if (stages == 2) {  // bin only: no tensors involved, but the stage offsets follow dout's layout
    if (bin_split) p.bin_stride = C;
    REQUIRE(workspace && aligned(workspace, 256), ARFE_E_ALIGN, "%s: workspace NULL or misaligned", fn);
    REQUIRE(workspace_bytes >= arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W), ARFE_E_SHAPE,
            "%s: workspace too small", fn);
    return cuda_result(arfe::launch_roi_fuse_backward_pull(p, dtype, workspace, workspace_bytes, 2, (cudaStream_t)stream), fn);
  }
  REQUIRE(plan_ready >= 0 && plan_ready <= 2, ARFE_E_ENUM, "%s: plan_ready must be 0, 1 or 2", fn);
  if (dout_regions) {
    rc = set_split_regions(fn, p, dout_regions, regions, C, dtype, &dout);
    if (rc) return rc;
  }
  REQUIRE(dout && workspace, ARFE_E_NULL, "%s: dout/workspace is NULL", fn);
  REQUIRE(aligned(dout, 16) && aligned(workspace, 256), ARFE_E_ALIGN, "%s: dout (16) / workspace (256) misaligned", fn);
  REQUIRE(workspace_bytes >= arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W), ARFE_E_SHAPE,
          "%s: workspace too small (%zu < %zu)", fn, workspace_bytes, arfe::roi_pull_workspace_bytes(K, regions, L, B, H, W));
  p.dout = dout; p.dout_cl = 1;
  p.debug_skip = ARFE_KNOB_ENV("ARFE_BWD_SKIP", 0);
  rc = cuda_result(arfe::launch_roi_fuse_backward_pull(p, dtype, workspace, workspace_bytes, stages, (cudaStream_t)stream), fn);
  if (rc) return rc;
  // regions whose tap tables did not fit the workspace records: atomic kernel, adds on top
  p.debug_skip = 0;
  p.flag_list = arfe::roi_pull_flag_list(K, regions, L, B, H, W, workspace, &p.flag_count);
  p.bwd_vec = 0;
  return cuda_result(arfe::launch_roi_fuse_backward(p, dtype, ARFE_NHWC, (cudaStream_t)stream), fn);
}

int arfe_roi_fuse_backward_pull(const void* dout, const int32_t* H, const int32_t* W,
                                const float* spatial_scale, int L, int B, int C, const float* rois,
                                int K, int regions, float facs, int PH, int PW, int sampling_ratio,
                                float finest_scale, int dtype, float* const* dfeats,
                                void* workspace, size_t workspace_bytes, int plan_ready, void* stream) {
  return backward_pull_impl("arfe_roi_fuse_backward_pull", dout, nullptr, H, W, spatial_scale, L, B, C,
                            rois, K, regions, facs, PH, PW, sampling_ratio, finest_scale, dtype, dfeats,
                            workspace, workspace_bytes, plan_ready, stream);
}

int arfe_roi_pull_bin(const int32_t* H, const int32_t* W, const float* spatial_scale, int L, int B, int C,
                      const float* rois, int K, int regions, float facs, int PH, int PW,
                      int sampling_ratio, float finest_scale, int dtype, int split_regions,
                      void* workspace, size_t workspace_bytes, void* stream) {
  return backward_pull_impl("arfe_roi_pull_bin", nullptr, nullptr, H, W, spatial_scale, L, B, C, rois, K,
                            regions, facs, PH, PW, sampling_ratio, finest_scale, dtype, nullptr, workspace,
                            workspace_bytes, split_regions ? kBinOnlySplit : kBinOnly, stream);
}

int arfe_roi_fuse_backward_pull_split(const void* const* dout_regions, const int32_t* H, const int32_t* W,
                                      const float* spatial_scale, int L, int B, int C,
                                      const float* rois, int K, int regions, float facs, int PH,
                                      int PW, int sampling_ratio, float finest_scale, int dtype,
                                      float* const* dfeats, void* workspace, size_t workspace_bytes,
                                      int plan_ready, void* stream) {
  if (!dout_regions && K > 0) return fail(ARFE_E_NULL, "arfe_roi_fuse_backward_pull_split: dout_regions is NULL");
  return backward_pull_impl("arfe_roi_fuse_backward_pull_split", nullptr, dout_regions, H, W, spatial_scale,
                            L, B, C, rois, K, regions, facs, PH, PW, sampling_ratio, finest_scale, dtype,
                            dfeats, workspace, workspace_bytes, plan_ready, stream);
}

int arfe_roi_align_forward(const void* input, const float* rois, float spatial_scale,
                           int pooled_height, int pooled_width, int sampling_ratio, int aligned_,
                           int B, int C, int H, int W, int K, int dtype, int layout,
                           void* output, void* stream) {
  REQUIRE(aligned_ == 1, ARFE_E_UNSUPPORTED,
          "arfe_roi_align_forward: aligned=False (legacy v1 RoIAlign) is not implemented");
  const void* feats[1] = {input};
  const int32_t h[1] = {H}, w[1] = {W};
  const float s[1] = {spatial_scale};
  return arfe_roi_fuse_forward(feats, h, w, s, 1, B, C, rois, K, 1, 1.0f, pooled_height,
                               pooled_width, sampling_ratio, 56.0f, dtype, layout, ARFE_NCHW, output,
                               nullptr, nullptr, stream);
}

int arfe_roi_align_backward(const void* grad, const float* rois, float spatial_scale,
                            int pooled_height, int pooled_width, int B, int C, int H, int W,
                            int K, int sampling_ratio, int aligned_, int dtype, int layout,
                            float* grad_input, void* stream) {
  REQUIRE(aligned_ == 1, ARFE_E_UNSUPPORTED,
          "arfe_roi_align_backward: aligned=False (legacy v1 RoIAlign) is not implemented");
  float* dfeats[1] = {grad_input};
  const int32_t h[1] = {H}, w[1] = {W};
  const float s[1] = {spatial_scale};
  return arfe_roi_fuse_backward(grad, ARFE_NCHW, h, w, s, 1, B, C, rois, K, 1, 1.0f, pooled_height,
                                pooled_width, sampling_ratio, 56.0f, dtype, layout, dfeats, stream);
}

int arfe_roi_fuse_taps(const int32_t* H, const int32_t* W, const float* spatial_scale, int L,
                       const float* rois, int K, int regions, float facs, int PH, int PW,
                       int sampling_ratio, float finest_scale, int max_grid, int32_t* lvl,
                       int32_t* grid, float* boxes, int32_t* ylo, int32_t* yhi, float* ywl,
                       float* ywh, int32_t* xlo, int32_t* xhi, float* xwl, float* xwh,
                       void* stream) {
  const char* fn = "arfe_roi_fuse_taps";
  arfe::RoiFuseParams p;
  int rc = fill_roi_params(fn, p, H, W, spatial_scale, L, 1, 1, rois, K, regions, facs, PH, PW,
                           sampling_ratio, finest_scale, ARFE_F32, ARFE_NCHW);
  if (rc) return rc;
  if (K == 0) return ARFE_OK;
  REQUIRE(max_grid >= 1, ARFE_E_SHAPE, "%s: max_grid=%d", fn, max_grid);
  DeviceGuard guard(rois);
  REQUIRE((ylo == nullptr) == (yhi == nullptr) && (ylo == nullptr) == (ywl == nullptr) &&
              (ylo == nullptr) == (ywh == nullptr) && (xlo == nullptr) == (xhi == nullptr) &&
              (xlo == nullptr) == (xwl == nullptr) && (xlo == nullptr) == (xwh == nullptr),
          ARFE_E_NULL, "%s: tap outputs must be given per axis as a complete set", fn);
  return cuda_result(arfe::launch_roi_fuse_taps(p, max_grid, lvl, grid, boxes, ylo, yhi, ywl, ywh,
                                                xlo, xhi, xwl, xwh, (cudaStream_t)stream), fn);
}

int arfe_rff_gate_forward(const void* ori, int64_t ori_roi_stride, const void* a, const void* b,
                          void* out, int64_t K, int64_t n_per_roi, int dtype, void* stream) {
  const char* fn = "arfe_rff_gate_forward";
  REQUIRE(dtype == ARFE_F32 || dtype == ARFE_BF16, ARFE_E_ENUM, "%s: unknown dtype %d", fn, dtype);
  REQUIRE(K >= 0 && n_per_roi >= 1 && ori_roi_stride >= n_per_roi, ARFE_E_SHAPE,
          "%s: bad K=%lld n=%lld stride=%lld", fn, (long long)K, (long long)n_per_roi, (long long)ori_roi_stride);
  if (K == 0) return ARFE_OK;
  REQUIRE(ori && a && b && out, ARFE_E_NULL, "%s: NULL tensor", fn);
  REQUIRE(n_per_roi < (1ll << 31), ARFE_E_SHAPE, "%s: n_per_roi too large", fn);
  DeviceGuard guard(out);
  return cuda_result(arfe::launch_rff_gate_forward(ori, ori_roi_stride, a, b, out, K, n_per_roi,
                                                   dtype, (cudaStream_t)stream), fn);
}

int arfe_rff_gate_backward(const void* g, const void* ori, int64_t ori_roi_stride, const void* a,
                           const void* b, void* d_ori, int64_t d_ori_roi_stride, void* d_ab,
                           int64_t K, int64_t n_per_roi, int dtype, void* stream) {
  const char* fn = "arfe_rff_gate_backward";
  REQUIRE(dtype == ARFE_F32 || dtype == ARFE_BF16, ARFE_E_ENUM, "%s: unknown dtype %d", fn, dtype);
  REQUIRE(K >= 0 && n_per_roi >= 1 && ori_roi_stride >= n_per_roi, ARFE_E_SHAPE,
          "%s: bad K=%lld n=%lld stride=%lld", fn, (long long)K, (long long)n_per_roi, (long long)ori_roi_stride);
  if (K == 0) return ARFE_OK;
  REQUIRE(g && ori && a && b && d_ori && d_ab, ARFE_E_NULL, "%s: NULL tensor", fn);
  REQUIRE(n_per_roi < (1ll << 31), ARFE_E_SHAPE, "%s: n_per_roi too large", fn);
  REQUIRE(d_ori_roi_stride >= n_per_roi, ARFE_E_SHAPE, "%s: d_ori stride < n_per_roi", fn);
  DeviceGuard guard(d_ab);
  return cuda_result(arfe::launch_rff_gate_backward(g, ori, ori_roi_stride, a, b, d_ori, d_ori_roi_stride, d_ab, K,
                                                    n_per_roi, dtype, (cudaStream_t)stream), fn);
}

static int softmax_fuse_impl(const char* fn, int backward, const void* const* regions, const int64_t* region_strides,
                             const void* logits, const int64_t* logit_strides, void* out, const void* dout,
                             const int64_t* out_strides, void* const* d_regions, void* d_logits, int64_t K, int PP,
                             int C, int dtype, void* stream) {
  REQUIRE(dtype == ARFE_F32 || dtype == ARFE_BF16, ARFE_E_ENUM, "%s: unknown dtype %d", fn, dtype);
  REQUIRE(K >= 0 && PP >= 1 && C >= 1, ARFE_E_SHAPE, "%s: bad K=%lld PP=%d C=%d", fn, (long long)K, PP, C);
  if (K == 0) return ARFE_OK;
  REQUIRE(regions && region_strides && logits && logit_strides && out_strides, ARFE_E_NULL, "%s: NULL argument", fn);
  REQUIRE(regions[0] && regions[1] && regions[2], ARFE_E_NULL, "%s: NULL region tensor", fn);
  if (backward) {
    REQUIRE(dout && d_regions && d_logits && d_regions[0] && d_regions[1] && d_regions[2], ARFE_E_NULL, "%s: NULL gradient tensor", fn);
  } else {
    REQUIRE(out, ARFE_E_NULL, "%s: out is NULL", fn);
  }
  for (int i = 0; i < 3; ++i)
    REQUIRE(region_strides[i] >= 1 && out_strides[i] >= 1 && logit_strides[i] >= 1, ARFE_E_SHAPE, "%s: strides must be >= 1", fn);
  DeviceGuard guard(logits);
  return cuda_result(arfe::launch_rff_softmax_fuse(backward, regions, region_strides, logits, logit_strides, out, dout,
                                                   out_strides, d_regions, d_logits, K, PP, C, dtype, (cudaStream_t)stream), fn);
}

int arfe_rff_softmax_fuse_forward(const void* const* regions, const int64_t* region_strides, const void* logits,
                                  const int64_t* logit_strides, void* out, const int64_t* out_strides, int64_t K,
                                  int PP, int C, int dtype, void* stream) {
  return softmax_fuse_impl("arfe_rff_softmax_fuse_forward", 0, regions, region_strides, logits, logit_strides, out,
                           nullptr, out_strides, nullptr, nullptr, K, PP, C, dtype, stream);
}

int arfe_rff_softmax_fuse_backward(const void* dout, const void* const* regions, const int64_t* region_strides,
                                   const void* logits, const int64_t* logit_strides, const int64_t* out_strides,
                                   void* const* d_regions, void* d_logits, int64_t K, int PP, int C, int dtype,
                                   void* stream) {
  return softmax_fuse_impl("arfe_rff_softmax_fuse_backward", 1, regions, region_strides, logits, logit_strides, nullptr,
                           dout, out_strides, d_regions, d_logits, K, PP, C, dtype, stream);
}

size_t arfe_nms_workspace_bytes(int n) { return n > 0 ? arfe::nms_workspace_bytes(n) : 0; }

int arfe_nms(const float* dets_sorted, int n, float iou_threshold, void* workspace, size_t workspace_bytes,
             int64_t* keep, int32_t* num_keep, void* stream) {
  const char* fn = "arfe_nms";
  REQUIRE(n >= 0, ARFE_E_SHAPE, "%s: n=%d", fn, n);
  REQUIRE(num_keep, ARFE_E_NULL, "%s: num_keep is NULL", fn);
  if (n > 0) {
    REQUIRE(dets_sorted && keep && workspace, ARFE_E_NULL, "%s: NULL argument", fn);
    REQUIRE(aligned(workspace, 8) && aligned(dets_sorted, 4), ARFE_E_ALIGN, "%s: workspace (8) / dets (4) misaligned", fn);
    REQUIRE(workspace_bytes >= arfe::nms_workspace_bytes(n), ARFE_E_SHAPE, "%s: workspace too small (%zu < %zu)", fn,
            workspace_bytes, arfe::nms_workspace_bytes(n));
  }
  DeviceGuard guard(num_keep);
  return cuda_result(arfe::launch_nms(dets_sorted, n, iou_threshold, workspace, keep, num_keep, (cudaStream_t)stream), fn);
}

int arfe_bbox2roi(const float* const* boxes, const int32_t* counts, const int32_t* cols, int B, float* rois,
                  void* stream) {
  const char* fn = "arfe_bbox2roi";
  REQUIRE(B >= 0, ARFE_E_SHAPE, "%s: B=%d", fn, B);
  if (B == 0) return ARFE_OK;
  REQUIRE(boxes && counts && cols, ARFE_E_NULL, "%s: NULL argument", fn);
  long long total = 0;
  for (int i = 0; i < B; ++i) {
    REQUIRE(counts[i] >= 0 && cols[i] >= 4, ARFE_E_SHAPE, "%s: list %d has %d rows of %d floats", fn, i, counts[i], cols[i]);
    REQUIRE(counts[i] == 0 || boxes[i], ARFE_E_NULL, "%s: boxes[%d] is NULL", fn, i);
    total += counts[i];
  }
  if (total == 0) return ARFE_OK;
  REQUIRE(rois, ARFE_E_NULL, "%s: rois is NULL", fn);
  REQUIRE(total < (1ll << 31), ARFE_E_SHAPE, "%s: too many boxes", fn);
  DeviceGuard guard(rois);
  return cuda_result(arfe::launch_bbox2roi(boxes, counts, cols, B, rois, (cudaStream_t)stream), fn);
}

size_t arfe_fpn_gate_conv_workspace_bytes(int L, int B, const int32_t* H, const int32_t* W) {
  if (L < 1 || L > ARFE_MAX_LEVELS || B < 0 || !H || !W) return 0;
  return arfe::fpn_gate_conv_workspace_bytes(L, B, H, W);
}

int arfe_fpn_gate_conv_forward(const void* const* feats, const float* const* w1, const float* const* b1,
                               const float* const* w2, const float* const* b2, const int32_t* H, const int32_t* W,
                               int L, int B, int C, int dtype, int layout, void* workspace, size_t workspace_bytes,
                               void* const* g1, void* const* g2, void* stream) {
  const char* fn = "arfe_fpn_gate_conv_forward";
  int rc = check_common(fn, L, B, C, H, W, dtype, layout);
  if (rc) return rc;
  REQUIRE(layout == ARFE_NHWC, ARFE_E_UNSUPPORTED, "%s: channels-last feature maps only", fn);
  if (B == 0) return ARFE_OK;
  REQUIRE(feats && w1 && b1 && w2 && b2 && g1 && g2 && workspace, ARFE_E_NULL, "%s: NULL argument", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(feats[l] && w1[l] && b1[l] && w2[l] && b2[l] && g1[l] && g2[l], ARFE_E_NULL, "%s: NULL tensor at level %d", fn, l);
    REQUIRE(aligned(feats[l], 16), ARFE_E_ALIGN, "%s: feats[%d] must be 16-byte aligned", fn, l);
  }
  REQUIRE(aligned(workspace, 16), ARFE_E_ALIGN, "%s: workspace must be 16-byte aligned", fn);
  REQUIRE(workspace_bytes >= arfe::fpn_gate_conv_workspace_bytes(L, B, H, W), ARFE_E_SHAPE, "%s: workspace too small", fn);
  DeviceGuard guard(workspace);
  const cudaError_t e = arfe::launch_fpn_gate_conv_forward(feats, w1, b1, w2, b2, H, W, L, B, C, dtype, workspace, g1,
                                                           g2, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported)
    return fail(ARFE_E_UNSUPPORTED, "%s: needs C %% %d == 0 and C <= %d", fn, dtype == ARFE_F32 ? 4 : 8,
                dtype == ARFE_F32 ? 256 : 512);
  return cuda_result(e, fn);
}

int arfe_nonlocal_default_split(int B, int HW) { return arfe::nonlocal_default_split(B, HW); }

size_t arfe_nonlocal_workspace_bytes(int B, int HW, int D, int nsplit) {
  if (B < 0 || HW < 1 || nsplit < 1 || nsplit > 8 || (D != 64 && D != 128 && D != 256)) return 0;
  return arfe::nonlocal_workspace_bytes(B, HW, D, nsplit);
}

int arfe_nonlocal_attention_forward(const void* theta, const void* phi, const void* g, void* y, int B, int HW, int D,
                                    int dtype, int layout, float scale, int nsplit, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  const char* fn = "arfe_nonlocal_attention_forward";
  REQUIRE(dtype == ARFE_F32 || dtype == ARFE_BF16, ARFE_E_UNSUPPORTED, "%s: dtype=%d", fn, dtype);
  REQUIRE(layout == ARFE_NCHW || layout == ARFE_NHWC, ARFE_E_UNSUPPORTED, "%s: layout=%d", fn, layout);
  REQUIRE(D == 64 || D == 128 || D == 256, ARFE_E_UNSUPPORTED, "%s: inter_channels must be 64, 128 or 256 (got %d)", fn, D);
  REQUIRE(B >= 0 && HW >= 1, ARFE_E_SHAPE, "%s: B=%d HW=%d", fn, B, HW);
  REQUIRE(scale > 0.f, ARFE_E_SHAPE, "%s: scale must be positive", fn);
  if (B == 0) return ARFE_OK;
  REQUIRE(B <= 21845, ARFE_E_SHAPE, "%s: B=%d exceeds the grid limit", fn, B);
  REQUIRE(nsplit >= 1 && nsplit <= 8 && nsplit <= (HW + 63) / 64, ARFE_E_SHAPE, "%s: nsplit=%d", fn, nsplit);
  REQUIRE(theta && phi && g && y && workspace, ARFE_E_NULL, "%s: NULL argument", fn);
  REQUIRE(aligned(workspace, 1024), ARFE_E_ALIGN, "%s: workspace must be 1024-byte aligned", fn);
  REQUIRE(aligned(theta, 16) && aligned(phi, 16) && aligned(g, 16) && aligned(y, 16), ARFE_E_ALIGN,
          "%s: theta, phi, g, y must be 16-byte aligned", fn);
  REQUIRE(workspace_bytes >= arfe::nonlocal_workspace_bytes(B, HW, D, nsplit), ARFE_E_SHAPE, "%s: workspace too small", fn);
  DeviceGuard guard(workspace);
  return cuda_result(arfe::launch_nonlocal_attention(theta, phi, g, y, B, HW, D, dtype, layout == ARFE_NHWC, scale,
                                                     workspace, nsplit, (cudaStream_t)stream), fn);
}

int arfe_fpn_gate_conv_backward(const void* const* feats, const float* const* w1, const float* const* w2,
                                const void* const* dg1, const void* const* dg2, const int32_t* H, const int32_t* W,
                                int L, int B, int C, int dtype, int layout, void* const* dx, float* const* dw1,
                                float* const* db1, float* const* dw2, float* const* db2, void* stream) {
  const char* fn = "arfe_fpn_gate_conv_backward";
  int rc = check_common(fn, L, B, C, H, W, dtype, layout);
  if (rc) return rc;
  REQUIRE(layout == ARFE_NHWC, ARFE_E_UNSUPPORTED, "%s: channels-last feature maps only", fn);
  if (B == 0) return ARFE_OK;
  REQUIRE(feats && w1 && w2 && dg1 && dg2 && dw1 && db1 && dw2 && db2, ARFE_E_NULL, "%s: NULL argument", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(feats[l] && w1[l] && w2[l] && dg1[l] && dg2[l] && dw1[l] && db1[l] && dw2[l] && db2[l] && (!dx || dx[l]),
            ARFE_E_NULL, "%s: NULL tensor at level %d", fn, l);
    REQUIRE(aligned(feats[l], 16) && (!dx || aligned(dx[l], 16)), ARFE_E_ALIGN, "%s: feats / dx[%d] must be 16-byte aligned", fn, l);
  }
  DeviceGuard guard(feats[0]);
  const cudaError_t e = arfe::launch_fpn_gate_conv_backward(feats, w1, w2, dg1, dg2, H, W, L, B, C, dtype, dx, dw1, db1,
                                                            dw2, db2, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported) return fail(ARFE_E_UNSUPPORTED, "%s: needs C %% 4 == 0 and C <= 512", fn);
  return cuda_result(e, fn);
}

int arfe_nonlocal_backward_rows(const float* S, const float* dP, void* P_bf16, void* dS_bf16, int64_t rows, int n,
                                float scale, void* stream) {
  const char* fn = "arfe_nonlocal_backward_rows";
  REQUIRE(rows >= 0 && n >= 1 && n <= 50 * 1024, ARFE_E_SHAPE, "%s: rows=%lld n=%d (n <= 51200)", fn, (long long)rows, n);
  if (rows == 0) return ARFE_OK;
  REQUIRE(S && dP && P_bf16 && dS_bf16, ARFE_E_NULL, "%s: NULL argument", fn);
  DeviceGuard guard(S);
  return cuda_result(arfe::launch_nonlocal_backward_rows(S, dP, P_bf16, dS_bf16, rows, n, scale, (cudaStream_t)stream), fn);
}

static int fill_fpn(const char* fn, arfe::FpnParams& p, const int32_t* H, const int32_t* W, int L,
                    int B, int C, int dtype, int layout) {
  int rc = check_common(fn, L, B, C, H, W, dtype, layout);
  if (rc) return rc;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < L; ++l) { p.H[l] = H[l]; p.W[l] = W[l]; }
  p.L = L; p.B = B; p.C = C;
  return ARFE_OK;
}

int arfe_fpn_gather_forward(const void* const* feats, const int32_t* H, const int32_t* W, int L,
                            int B, int C, int refine_level, int dtype, int layout, void* out,
                            uint8_t* argmax, void* stream) {
  const char* fn = "arfe_fpn_gather_forward";
  arfe::FpnParams p;
  int rc = fill_fpn(fn, p, H, W, L, B, C, dtype, layout);
  if (rc) return rc;
  REQUIRE(refine_level >= 0 && refine_level < L, ARFE_E_SHAPE, "%s: refine_level=%d", fn, refine_level);
  if (B == 0) return ARFE_OK;
  REQUIRE(feats && out, ARFE_E_NULL, "%s: feats/out is NULL", fn);
  p.refine_level = refine_level; p.Hr = H[refine_level]; p.Wr = W[refine_level];
  for (int l = 0; l < L; ++l) {
    REQUIRE(feats[l], ARFE_E_NULL, "%s: feats[%d] is NULL", fn, l);
    p.feats[l] = feats[l];
    if (l < refine_level) {
      const long long kh = (H[l] + p.Hr - 1) / p.Hr + 1, kw = (W[l] + p.Wr - 1) / p.Wr + 1;
      REQUIRE(argmax == nullptr || kh * kw <= 255, ARFE_E_UNSUPPORTED,
              "%s: pooling window of level %d too large for uint8 argmax", fn, l);
    }
  }
  p.gathered = out; p.argmax = argmax;
  DeviceGuard guard(out);
  return cuda_result(arfe::launch_fpn_gather_forward(p, dtype, layout, (cudaStream_t)stream), fn);
}

int arfe_fpn_gather_backward_acc(const void* dout, const uint8_t* argmax, const int32_t* H,
                                 const int32_t* W, int L, int B, int C, int refine_level, int dtype,
                                 int layout, const float* const* addend, void* const* dfeats,
                                 void* stream) {
  const char* fn = "arfe_fpn_gather_backward";
  arfe::FpnParams p;
  int rc = fill_fpn(fn, p, H, W, L, B, C, dtype, layout);
  if (rc) return rc;
  REQUIRE(refine_level >= 0 && refine_level < L, ARFE_E_SHAPE, "%s: refine_level=%d", fn, refine_level);
  if (B == 0) return ARFE_OK;
  REQUIRE(dout && dfeats, ARFE_E_NULL, "%s: dout/dfeats is NULL", fn);
  REQUIRE(refine_level == 0 || argmax, ARFE_E_NULL, "%s: argmax is NULL", fn);
  p.refine_level = refine_level; p.Hr = H[refine_level]; p.Wr = W[refine_level];
  for (int l = 0; l < L; ++l) {
    REQUIRE(dfeats[l], ARFE_E_NULL, "%s: dfeats[%d] is NULL", fn, l);
    p.outs[l] = dfeats[l];
    if (addend && addend[l]) {
      REQUIRE(aligned(addend[l], 16), ARFE_E_ALIGN, "%s: addend[%d] must be 16-byte aligned", fn, l);
      p.addend[l] = addend[l];
    }
  }
  p.gathered = const_cast<void*>(dout); p.argmax = const_cast<uint8_t*>(argmax);
  DeviceGuard guard(dout);
  return cuda_result(arfe::launch_fpn_gather_backward(p, dtype, layout, (cudaStream_t)stream), fn);
}

int arfe_fpn_gather_backward(const void* dout, const uint8_t* argmax, const int32_t* H,
                             const int32_t* W, int L, int B, int C, int refine_level, int dtype,
                             int layout, void* const* dfeats, void* stream) {
  return arfe_fpn_gather_backward_acc(dout, argmax, H, W, L, B, C, refine_level, dtype, layout, nullptr,
                                      dfeats, stream);
}

int arfe_fpn_apply_forward(const void* const* feats, const void* bsf, const void* const* g1,
                           const void* const* g2, const int32_t* H, const int32_t* W, int L, int B,
                           int C, int Hr, int Wr, int dtype, int layout, void* const* outs,
                           void* stream) {
  const char* fn = "arfe_fpn_apply_forward";
  arfe::FpnParams p;
  int rc = fill_fpn(fn, p, H, W, L, B, C, dtype, layout);
  if (rc) return rc;
  REQUIRE(Hr >= 1 && Wr >= 1, ARFE_E_SHAPE, "%s: bsf size %dx%d", fn, Hr, Wr);
  if (B == 0) return ARFE_OK;
  REQUIRE(feats && bsf && g1 && g2 && outs, ARFE_E_NULL, "%s: NULL argument", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(feats[l] && g1[l] && g2[l] && outs[l], ARFE_E_NULL, "%s: NULL tensor at level %d", fn, l);
    p.feats[l] = feats[l]; p.g1[l] = g1[l]; p.g2[l] = g2[l]; p.outs[l] = outs[l];
  }
  p.bsf = bsf; p.Hr = Hr; p.Wr = Wr;
  DeviceGuard guard(bsf);
  return cuda_result(arfe::launch_fpn_apply_forward(p, dtype, layout, (cudaStream_t)stream), fn);
}

int arfe_fpn_apply_backward(const void* const* douts, const void* bsf, const void* const* g1,
                            const void* const* g2, const int32_t* H, const int32_t* W, int L,
                            int B, int C, int Hr, int Wr, int dtype, int layout, float* dbsf,
                            float* const* dg1, float* const* dg2, void* stream) {
  const char* fn = "arfe_fpn_apply_backward";
  arfe::FpnParams p;
  int rc = fill_fpn(fn, p, H, W, L, B, C, dtype, layout);
  if (rc) return rc;
  REQUIRE(Hr >= 1 && Wr >= 1, ARFE_E_SHAPE, "%s: bsf size %dx%d", fn, Hr, Wr);
  if (B == 0) return ARFE_OK;
  REQUIRE(douts && bsf && g1 && g2 && dbsf && dg1 && dg2, ARFE_E_NULL, "%s: NULL argument", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(douts[l] && g1[l] && g2[l] && dg1[l] && dg2[l], ARFE_E_NULL, "%s: NULL tensor at level %d", fn, l);
    p.feats[l] = douts[l]; p.g1[l] = g1[l]; p.g2[l] = g2[l]; p.dg1[l] = dg1[l]; p.dg2[l] = dg2[l];
  }
  p.bsf = bsf; p.Hr = Hr; p.Wr = Wr; p.dbsf = dbsf;
  DeviceGuard guard(dbsf);
  return cuda_result(arfe::launch_fpn_apply_backward(p, dtype, layout, (cudaStream_t)stream), fn);
}

int arfe_fpn_backward_fused(const void* const* douts, int douts_f32, const void* bsf, const void* const* g1,
                            const void* const* g2, const void* dgathered, const uint8_t* argmax,
                            const int32_t* H, const int32_t* W, int L, int B, int C, int refine_level,
                            int dtype, int layout, float* dbsf, float* const* dg1, float* const* dg2,
                            void* const* dx, void* stream) {
  const char* fn = "arfe_fpn_backward_fused";
  arfe::FpnParams p;
  int rc = fill_fpn(fn, p, H, W, L, B, C, dtype, layout);
  if (rc) return rc;
  REQUIRE(refine_level >= 0 && refine_level < L, ARFE_E_SHAPE, "%s: refine_level=%d", fn, refine_level);
  REQUIRE(layout == ARFE_NHWC, ARFE_E_UNSUPPORTED,
          "%s: channels-last tensors only (use arfe_fpn_apply_backward + arfe_fpn_gather_backward_acc)", fn);
  if (B == 0) return ARFE_OK;
  REQUIRE(douts && bsf && g1 && g2 && dgathered && dbsf && dg1 && dg2 && dx, ARFE_E_NULL, "%s: NULL argument", fn);
  REQUIRE(refine_level == 0 || argmax, ARFE_E_NULL, "%s: argmax is NULL", fn);
  for (int l = 0; l < L; ++l) {
    REQUIRE(douts[l] && g1[l] && g2[l] && dg1[l] && dg2[l] && dx[l], ARFE_E_NULL, "%s: NULL tensor at level %d", fn, l);
    p.feats[l] = douts[l]; p.g1[l] = g1[l]; p.g2[l] = g2[l]; p.dg1[l] = dg1[l]; p.dg2[l] = dg2[l];
    p.outs[l] = dx[l];
  }
  p.refine_level = refine_level; p.Hr = H[refine_level]; p.Wr = W[refine_level];
  p.bsf = bsf; p.dbsf = dbsf; p.gathered = const_cast<void*>(dgathered); p.argmax = const_cast<uint8_t*>(argmax);
  DeviceGuard guard(dbsf);
  const cudaError_t e = arfe::launch_fpn_backward_fused_cl(p, dtype, douts_f32 ? 1 : 0, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported)
    return fail(ARFE_E_UNSUPPORTED,
                "%s: needs C %% %d == 0, C <= %d, 16-byte aligned tensors and integer pooling ratios below the "
                "refine level (use arfe_fpn_apply_backward + arfe_fpn_gather_backward_acc)",
                fn, dtype == ARFE_F32 ? 4 : 8, dtype == ARFE_F32 ? 256 : 512);
  return cuda_result(e, fn);
}

}  // extern "C"
