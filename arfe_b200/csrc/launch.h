// Internal launcher interface between capi.cu (argument validation, C ABI) and
// the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef ARFE_PROFILE
#include <stdlib.h>
#endif

// Profiling knobs (phase skipping, environment switches) exist only in the
// -DARFE_PROFILE build (libarfe_b200_prof.so, used by scripts/*knobs*.py).  The
// shipped library contains neither the skip branches nor a getenv: both macros
// fold to constants.
#ifdef ARFE_PROFILE
#define ARFE_SKIP(p, bits) (((p).debug_skip & (bits)) != 0)
#define ARFE_KNOB_ENV(name, dflt) ([] { const char* ev_ = getenv(name); return ev_ ? atoi(ev_) : (dflt); }())
#else
#define ARFE_SKIP(p, bits) false
#define ARFE_KNOB_ENV(name, dflt) (dflt)
#endif

namespace arfe {

constexpr int kMaxLevels = 8;  // == ARFE_MAX_LEVELS
constexpr int kMaxPool = 32;   // == ARFE_MAX_POOL

// SM count of the CURRENT device (persistent grids are sized from it); cached per device.
inline int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n;
  }
  return cache[dev];
}

struct RoiFuseParams {
  const void* feats[kMaxLevels];  // forward: pyramid
  float* dfeats[kMaxLevels];      // backward: gradient pyramid (accumulated)
  int H[kMaxLevels], W[kMaxLevels];
  float scale[kMaxLevels];
  int L, B, C, K, R, PH, PW, sampling_ratio;
  float facs, finest_scale;
  const float* rois;
  void* out;         // forward output
  const void* dout;  // backward input
  int32_t* lvl_out;
  float* boxes_out;
  int bwd_vec;       // backward: 128-bit vector reductions (1) or scalar (0)
  int out_cl;        // forward: output channels-last [K][PH*PW][R*C] (1) or NCHW (0)
  int dout_cl;       // backward: dout channels-last (1) or NCHW (0)
  const int* flag_list;   // backward (atomic kernel as the pull fallback): region ids to process ...
  const int* flag_count;  // ... and how many (device memory)
  long long reg_off[3];  // channels-last out / dout: element offset of region r's block from the base pointer ...
  int bin_stride;        // ... and elements between consecutive bins (concatenated: r * C, R * C)
  int debug_skip;    // ARFE_PROFILE builds only (ARFE_FWD_SKIP / ARFE_BWD_SKIP): 1 compute, 2 staging, 4 write-out, 8 all but setup
};

cudaError_t launch_roi_fuse_forward(const RoiFuseParams& p, int dtype, int layout,
                                    cudaStream_t stream);
cudaError_t launch_roi_fuse_backward(const RoiFuseParams& p, int dtype, int layout,
                                     cudaStream_t stream);
cudaError_t launch_roi_fuse_forward_cl(const RoiFuseParams& p, int dtype, int out_cl,
                                       cudaStream_t stream);
size_t roi_pull_workspace_bytes(int K, int R, int L, int B, const int* H, const int* W);
cudaError_t launch_roi_fuse_backward_pull(const RoiFuseParams& p, int dtype, void* workspace,
                                          size_t workspace_bytes, int stages, cudaStream_t stream);
cudaError_t launch_roi_fuse_forward_plan(const RoiFuseParams& p, int dtype, void* workspace,
                                         size_t workspace_bytes, int stages, cudaStream_t stream);
const int* roi_pull_flag_list(int K, int R, int L, int B, const int* H, const int* W, void* workspace,
                              const int** count);
cudaError_t launch_roi_fuse_taps(const RoiFuseParams& p, int max_grid, int32_t* lvl,
                                 int32_t* grid, float* boxes, int32_t* ylo,
                                 int32_t* yhi, float* ywl, float* ywh, int32_t* xlo,
                                 int32_t* xhi, float* xwl, float* xwh,
                                 cudaStream_t stream);

cudaError_t launch_rff_gate_forward(const void* ori, int64_t ori_stride,
                                    const void* a, const void* b, void* out,
                                    int64_t K, int64_t n, int dtype,
                                    cudaStream_t stream);
cudaError_t launch_rff_gate_backward(const void* g, const void* ori,
                                     int64_t ori_stride, const void* a,
                                     const void* b, void* d_ori, int64_t d_ori_stride,
                                     void* d_ab, int64_t K, int64_t n, int dtype,
                                     cudaStream_t stream);

// softmax-over-regions fusion (rff_gate.cu); strides in elements: regions / out (k, bin, channel), logits (k, region, bin)
cudaError_t launch_rff_softmax_fuse(int backward, const void* const* reg, const int64_t* rstr, const void* logits,
                                    const int64_t* lstr, void* out, const void* dout, const int64_t* ostr,
                                    void* const* dreg, void* dlogits, int64_t K, int PP, int C, int dtype,
                                    cudaStream_t stream);

// AR-FPN gate convolutions, both 3x3 C -> 1 filters of every level in one pass over x (fpn_gate_conv.cu)
size_t fpn_gate_conv_workspace_bytes(int L, int B, const int* H, const int* W);
cudaError_t launch_fpn_gate_conv_forward(const void* const* feats, const float* const* w1, const float* const* b1,
                                         const float* const* w2, const float* const* b2, const int* H, const int* W,
                                         int L, int B, int C, int dtype, void* workspace, void* const* g1,
                                         void* const* g2, cudaStream_t stream);

cudaError_t launch_fpn_gate_conv_backward(const void* const* feats, const float* const* w1, const float* const* w2,
                                          const void* const* dg1, const void* const* dg2, const int* H, const int* W,
                                          int L, int B, int C, int dtype, void* const* dx, float* const* dw1,
                                          float* const* db1, float* const* dw2, float* const* db2,
                                          cudaStream_t stream);

// proposal side (proposals.cu)
size_t nms_workspace_bytes(int n);
cudaError_t launch_nms(const float* dets_sorted, int n, float thr, void* workspace, int64_t* keep, int* num_keep,
                       cudaStream_t stream);
cudaError_t launch_bbox2roi(const float* const* boxes, const int* counts, const int* cols, int B, float* rois,
                            cudaStream_t stream);

// NonLocal2D refine as a fused tensor-core attention (nonlocal_attn.cu)
int nonlocal_default_split(int B, int HW);
size_t nonlocal_workspace_bytes(int B, int HW, int D, int nsplit);
cudaError_t launch_nonlocal_attention(const void* theta, const void* phi, const void* g, void* y, int B, int HW, int D,
                                      int dtype, int in_cl, float scale, void* workspace, int nsplit,
                                      cudaStream_t stream);

cudaError_t launch_nonlocal_backward_rows(const float* S, const float* dP, void* Pb, void* dSb, long long rows, int n,
                                          float scale, cudaStream_t stream);

struct FpnParams {
  const void* feats[kMaxLevels];  // x_l (gather fwd / apply fwd) or dout_l (apply bwd)
  void* outs[kMaxLevels];         // out_l (apply fwd) or dx_l (gather bwd)
  const void* g1[kMaxLevels];
  const void* g2[kMaxLevels];
  float* dg1[kMaxLevels];
  float* dg2[kMaxLevels];
  int H[kMaxLevels], W[kMaxLevels];
  int L, B, C, refine_level, Hr, Wr;
  const void* bsf;        // apply: [B,C,Hr,Wr]
  void* gathered;         // gather fwd out / gather bwd dout
  uint8_t* argmax;        // gather
  float* dbsf;            // apply bwd
  const float* addend[kMaxLevels];  // gather bwd: fp32 tensor added to level l's gradient (NULL: none)
};

bool fpn_cl_ok(const FpnParams& p, int dtype, bool need_feats, bool need_outs);
cudaError_t launch_fpn_gather_forward_cl(const FpnParams& p, int dtype, cudaStream_t stream);
cudaError_t launch_fpn_gather_backward_cl(const FpnParams& p, int dtype, unsigned* mask,
                                          cudaStream_t stream);
cudaError_t launch_fpn_apply_forward_cl(const FpnParams& p, int dtype, cudaStream_t stream);
cudaError_t launch_fpn_apply_backward_cl(const FpnParams& p, int dtype, cudaStream_t stream);
cudaError_t launch_fpn_backward_fused_cl(const FpnParams& p, int dtype, int dout_f32, cudaStream_t stream);

cudaError_t launch_fpn_gather_forward(const FpnParams& p, int dtype, int layout,
                                      cudaStream_t stream);
cudaError_t launch_fpn_gather_backward(const FpnParams& p, int dtype, int layout,
                                       cudaStream_t stream);
cudaError_t launch_fpn_apply_forward(const FpnParams& p, int dtype, int layout,
                                     cudaStream_t stream);
cudaError_t launch_fpn_apply_backward(const FpnParams& p, int dtype, int layout,
                                      cudaStream_t stream);

}  // namespace arfe
