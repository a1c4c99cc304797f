/*
 * arfe_b200.h -- C ABI of libarfe_b200.so, the B200 (sm_100a) implementation of
 * ARFE's region-aware feature path (AR-FPN aggregation + AR-RFF RoI fusion).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point names the reference interface it replaces; paths are
 * relative to the reference tree (Fanzhongjie/ARFE, an mmdetection 2.0 fork).
 *
 * Conventions:
 *  - Every pointer marked "device" is a CUDA device pointer valid on the
 *    current device.  The CALLER allocates every buffer; the library never
 *    allocates, frees or retains a pointer.
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*) and
 *    the call returns immediately; nothing synchronises the host
 *    (the reference blocks with cudaDeviceSynchronize() after every forward,
 *    mmdet/ops/roi_align/src/cuda/roi_align_kernel_v2.cu:304).
 *  - Return value: 0 = ok; negative = argument error (ARFE_E_*); positive =
 *    the cudaError_t reported by the launch.  arfe_last_error() returns a
 *    thread-local human-readable message for the last non-zero return.
 *  - There is no CPU implementation behind these symbols.
 *  - Tensors are dense and contiguous in the stated layout.
 *      dtype : ARFE_F32 (float) or ARFE_BF16 (__nv_bfloat16); arithmetic and
 *              accumulation are always fp32, box/level/index math is exact
 *              fp32 in the reference's operation order.
 *      layout: ARFE_NCHW (reference layout) or ARFE_NHWC (torch channels_last
 *              memory format of the same logical NCHW tensor).
 *  - RoIs are fp32 [K,5] = (batch_index, x1, y1, x2, y2) in image pixels
 *    (mmdet/core/bbox/transforms.py:41-60).
 */
#ifndef ARFE_B200_H_
#define ARFE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARFE_VERSION 100 /* 0.1.0 */
#define ARFE_MAX_LEVELS 8
#define ARFE_MAX_POOL 32 /* max pooled height / width */

enum { ARFE_F32 = 0, ARFE_BF16 = 1 };
enum { ARFE_NCHW = 0, ARFE_NHWC = 1 };

enum {
  ARFE_OK = 0,
  ARFE_E_NULL = -1,   /* required pointer is NULL */
  ARFE_E_SHAPE = -2,  /* negative / zero / too large dimension */
  ARFE_E_ENUM = -3,   /* unknown dtype / layout / regions value */
  ARFE_E_ALIGN = -4,  /* pointer not aligned for the element type */
  ARFE_E_UNSUPPORTED = -5
};

int arfe_version(void);
const char* arfe_last_error(void);

/* ------------------------------------------------------------------------
 * AR-RFF extraction: fused region generation + RoI->level map + multi-level
 * RoIAlign, written straight into the concatenated tensor.
 *
 * Replaces, in one launch:
 *   get_adaptive_scale_rois(rois, facs)        mmdet/models/utils/additional.py:38-71
 *   SingleRoIExtractor.map_roi_levels          mmdet/models/roi_heads/roi_extractors/single_level.py:53-93
 *   SingleRoIExtractor.forward  (x regions)    .../single_level.py:109-152
 *   RoIAlign / roi_align_ext.forward_v2        mmdet/ops/roi_align/src/roi_align_ext.cpp:126-142
 *   torch.cat([ori, lw, lh], dim=1)            mmdet/models/roi_heads/standard_roi_head.py:138-155
 *
 * regions = 1: out[K, C, PH, PW]      (plain SingleRoIExtractor.forward)
 * regions = 3: out[K, 3*C, PH, PW]    channels [0,C) original box, [C,2C)
 *              adaptive_w ("lw"), [2C,3C) adaptive_h ("lh").
 * L = 1 skips the level map (single_level.py:120-123).
 * feats[l]: device, [B, C, H[l], W[l]] in `layout`; out: device, dense, in
 * `out_layout`: ARFE_NCHW [K][regions*C][PH][PW] (the reference's) or ARFE_NHWC
 * = torch channels_last of the same logical tensor, [K][PH][PW][regions*C]
 * (needs layout == ARFE_NHWC and C a multiple of 4 (fp32) / 8 (bf16)).
 * lvl_out  (optional, device int32 [regions, K]): level chosen per region box
 *          (-1: box whose scale is NaN, matches no level -> zero output row).
 * boxes_out(optional, device fp32 [regions, K, 5]): the region boxes used.
 * ---------------------------------------------------------------------- */
int arfe_roi_fuse_forward(const void* const* feats, const int32_t* H,
                          const int32_t* W, const float* spatial_scale, int L,
                          int B, int C, const float* rois, int K, int regions,
                          float facs, int PH, int PW, int sampling_ratio,
                          float finest_scale, int dtype, int layout,
                          int out_layout, void* out, int32_t* lvl_out,
                          float* boxes_out, void* stream);

/* Backward of the above w.r.t. the pyramid (no gradient for rois, like
 * RoIAlignFunction.backward, mmdet/ops/roi_align/roi_align.py:44-73, and
 * roi_align_ext.backward_v2, roi_align_ext.cpp:144-161).
 * dout: device [K, regions*C, PH, PW] dense in `dout_layout`, `dtype`.
 * dfeats[l]: device fp32 [B, C, H[l], W[l]] in `layout`; the kernel ADDS into
 * them (the caller zero-fills, as the reference's at::zeros does,
 * roi_align_kernel_v2.cu:325-326; one buffer per level receives the sum over
 * all regions, replacing the reference's 15 dense per-level/region grads). */
int arfe_roi_fuse_backward(const void* dout, int dout_layout, const int32_t* H,
                           const int32_t* W, const float* spatial_scale, int L,
                           int B, int C, const float* rois, int K, int regions,
                           float facs, int PH, int PW, int sampling_ratio,
                           float finest_scale, int dtype, int layout,
                           float* const* dfeats, void* stream);

/* The RoI plan.  arfe_roi_fuse_forward_plan and arfe_roi_fuse_backward_pull share
 * a caller-provided device workspace of arfe_roi_plan_bytes() bytes (256-byte
 * aligned, owned by the caller, reusable after the stream passes the last call
 * that reads it).  The forward writes the per-region tables (boxes, levels,
 * sampling windows, aggregated bilinear weights) into it and consumes them; a
 * backward for the SAME rois / shapes / pool size may reuse them
 * (plan_ready = 1) instead of rebuilding them (plan_ready = 0); plan_ready = 2:
 * the backward's tile bins are ready too (arfe_roi_pull_bin).
 * arfe_roi_fuse_pull_workspace_bytes is the same number (older name). */
size_t arfe_roi_plan_bytes(int K, int regions, int L, int B, const int32_t* H, const int32_t* W);

/* The plan depends only on the RoIs and the shapes, the tile binning of the
 * backward only on the plan: both may be issued early, on another stream, to
 * overlap unrelated work (the neck's kernels, the head's backward).
 *   arfe_roi_plan_build : writes the plan; afterwards the forward may be called with
 *                         plan_ready = 1 and the backward with plan_ready = 1.
 *   arfe_roi_pull_bin   : needs the plan; afterwards arfe_roi_fuse_backward_pull(_split)
 *                         may be called with plan_ready = 2 (plan and bins ready).
 * `dtype` must be the dtype of the tensors the plan will be used with (it bounds
 * the window width the forward's ring takes).  The caller orders the streams. */
int arfe_roi_plan_build(const int32_t* H, const int32_t* W, const float* spatial_scale, int L,
                        int B, int C, const float* rois, int K, int regions, float facs,
                        int PH, int PW, int sampling_ratio, float finest_scale, int dtype,
                        void* workspace, size_t workspace_bytes, void* stream);
int arfe_roi_pull_bin(const int32_t* H, const int32_t* W, const float* spatial_scale, int L,
                      int B, int C, const float* rois, int K, int regions, float facs,
                      int PH, int PW, int sampling_ratio, float finest_scale, int dtype,
                      int split_regions /* the backward will be the _split entry point */,
                      void* workspace, size_t workspace_bytes, void* stream);
size_t arfe_roi_fuse_pull_workspace_bytes(int K, int regions, int L, int B,
                                          const int32_t* H, const int32_t* W);

/* Forward for channels-last tensors (feats ARFE_NHWC, out [K][PH][PW][regions*C]),
 * same result as arfe_roi_fuse_forward; C a multiple of 4 (fp32) / 8 (bf16).
 * One launch builds the plan, a persistent kernel streams every region's
 * sampling window through shared memory with bulk async copies (each window
 * byte is fetched once), a third serves the few regions that do not fit. */
int arfe_roi_fuse_forward_plan(const void* const* feats, const int32_t* H, const int32_t* W,
                               const float* spatial_scale, int L, int B, int C,
                               const float* rois, int K, int regions, float facs,
                               int PH, int PW, int sampling_ratio, float finest_scale,
                               int dtype, void* out, int32_t* lvl_out, float* boxes_out,
                               void* workspace, size_t workspace_bytes, int plan_ready,
                               void* stream);

/* Split layout: the regions as separate tensors [K][PH][PW][C] (channels-last of
 * [K, C, PH, PW]) instead of one concatenated [K][PH][PW][regions*C] -- what the
 * head's convolutions and their backward produce / consume, so no torch.cat /
 * slice copies are needed around the kernels.  out_regions / dout_regions:
 * `regions` device pointers, 16-byte aligned, anywhere in memory.  Everything
 * else as in arfe_roi_fuse_forward_plan / arfe_roi_fuse_backward_pull. */
int arfe_roi_fuse_forward_plan_split(const void* const* feats, const int32_t* H, const int32_t* W,
                                     const float* spatial_scale, int L, int B, int C,
                                     const float* rois, int K, int regions, float facs,
                                     int PH, int PW, int sampling_ratio, float finest_scale,
                                     int dtype, void* const* out_regions, void* workspace,
                                     size_t workspace_bytes, int plan_ready, void* stream);
int arfe_roi_fuse_backward_pull_split(const void* const* dout_regions, const int32_t* H,
                                      const int32_t* W, const float* spatial_scale, int L, int B,
                                      int C, const float* rois, int K, int regions, float facs,
                                      int PH, int PW, int sampling_ratio, float finest_scale,
                                      int dtype, float* const* dfeats, void* workspace,
                                      size_t workspace_bytes, int plan_ready, void* stream);

/* Atomic-free backward for channels-last tensors ("pull"): every element of
 * every dfeats[l] (fp32, ARFE_NHWC) is WRITTEN exactly once, in a fixed
 * summation order (deterministic, unlike the reference's atomicAdd,
 * roi_align_kernel_v2.cu:251-258); no zero-fill by the caller.
 * dout: [K][PH][PW][regions*C] (ARFE_NHWC), `dtype`; C a multiple of 4 / 8.
 * workspace: see "The RoI plan" above. */
int arfe_roi_fuse_backward_pull(const void* dout, const int32_t* H, const int32_t* W,
                                const float* spatial_scale, int L, int B, int C,
                                const float* rois, int K, int regions, float facs,
                                int PH, int PW, int sampling_ratio,
                                float finest_scale, int dtype,
                                float* const* dfeats, void* workspace,
                                size_t workspace_bytes, int plan_ready, void* stream);

/* Operator-level twins of roi_align_ext.forward_v2 / backward_v2
 * (roi_align_ext.cpp:126-161; aligned=True only -- aligned=False is the legacy
 * v1 path, which is not built).  input [B,C,H,W], rois [K,5],
 * output [K,C,PH,PW]; grad_input is ADDED into (caller zero-fills). */
int arfe_roi_align_forward(const void* input, const float* rois,
                           float spatial_scale, int pooled_height,
                           int pooled_width, int sampling_ratio, int aligned,
                           int B, int C, int H, int W, int K, int dtype,
                           int layout, void* output, void* stream);
int arfe_roi_align_backward(const void* grad, const float* rois,
                            float spatial_scale, int pooled_height,
                            int pooled_width, int B, int C, int H, int W,
                            int K, int sampling_ratio, int aligned, int dtype,
                            int layout, float* grad_input, void* stream);

/* Parity instrumentation: the region boxes, levels, sampling grids and the
 * bilinear rows/columns/weights the kernels above use, computed by the same
 * device functions.  Shapes (R = regions):
 *   lvl [R,K] int32, grid [R,K,2] int32 (grid_h, grid_w), boxes [R,K,5] fp32,
 *   ylo/yhi [R,K,PH,max_grid] int32, ywl/ywh same shape fp32,
 *   xlo/xhi [R,K,PW,max_grid] int32, xwl/xwh same shape fp32.
 * Slots i >= grid are -2; samples outside the map are -1 (weights 0). */
int arfe_roi_fuse_taps(const int32_t* H, const int32_t* W,
                       const float* spatial_scale, int L, const float* rois,
                       int K, int regions, float facs, int PH, int PW,
                       int sampling_ratio, float finest_scale, int max_grid,
                       int32_t* lvl, int32_t* grid, float* boxes, int32_t* ylo,
                       int32_t* yhi, float* ywl, float* ywh, int32_t* xlo,
                       int32_t* xhi, float* xwl, float* xwh, void* stream);

/* ------------------------------------------------------------------------
 * AR-RFF fusion gate: out = ori + ori*(a + b) = ori*(1 + a + b)
 *   MultiBBoxHead.forward, mmdet/models/roi_heads/bbox_heads/multirois_bbox_head.py:175,182
 * ori is read in place from the concatenated tensor: roi k starts at
 * ori + k*ori_roi_stride elements and holds n_per_roi (= C*PH*PW) elements.
 * a, b, out: dense [K, n_per_roi].
 * backward: d_ori = g*(1+a+b), rows d_ori_roi_stride elements apart (so it can
 * be written in place into the gradient of the concatenated tensor: no cat of
 * the ori block), da = db = g*ori (one dense buffer, d_ab). */
int arfe_rff_gate_forward(const void* ori, int64_t ori_roi_stride, const void* a,
                          const void* b, void* out, int64_t K,
                          int64_t n_per_roi, int dtype, void* stream);
int arfe_rff_gate_backward(const void* g, const void* ori,
                           int64_t ori_roi_stride, const void* a, const void* b,
                           void* d_ori, int64_t d_ori_roi_stride, void* d_ab,
                           int64_t K, int64_t n_per_roi, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * AR-RFF, softmax-over-regions fusion -- the fusion of the paper's figure; in the
 * reference it is the commented block
 * mmdet/models/roi_heads/bbox_heads/multirois_bbox_head.py:187-197:
 *   ws  = softmax(logits, dim=1)                      logits [K, 3, PH, PW]
 *   out = r0 * ws[:,0] + r1 * ws[:,1] + r2 * ws[:,2]  r_j, out [K, C, PH, PW]
 * regions[3]: device pointers; element (k, bin, c) of every region tensor sits
 * at k*region_strides[0] + bin*region_strides[1] + c*region_strides[2] (elements)
 * -- one triple for the three, which covers NCHW tensors, channels-last tensors
 * and the channel slices of the concatenated [K, 3C, PH, PW] tensor alike; out /
 * dout / d_regions use out_strides the same way; logits / d_logits element
 * (k, j, bin) sits at k*logit_strides[0] + j*logit_strides[1] + bin*logit_strides[2].
 * PP = PH*PW.  Channels-last tensors (channel stride 1) take a warp-per-bin
 * kernel with 128-bit accesses; anything else a thread-per-bin kernel.
 * backward: d_regions[j] = dout * ws_j (written), d_logits_j = ws_j * (s_j -
 * sum_i ws_i s_i), s_j = sum_c dout * r_j (written). */
int arfe_rff_softmax_fuse_forward(const void* const* regions,
                                  const int64_t* region_strides,
                                  const void* logits,
                                  const int64_t* logit_strides, void* out,
                                  const int64_t* out_strides, int64_t K, int PP,
                                  int C, int dtype, void* stream);
int arfe_rff_softmax_fuse_backward(const void* dout, const void* const* regions,
                                   const int64_t* region_strides,
                                   const void* logits,
                                   const int64_t* logit_strides,
                                   const int64_t* out_strides,
                                   void* const* d_regions, void* d_logits,
                                   int64_t K, int PP, int C, int dtype,
                                   void* stream);

/* ------------------------------------------------------------------------
 * AR-FPN gather: every level resized to the refine level and averaged.
 *   WFPNDualSpatial.forward, mmdet/models/necks/wfpn_dual_spatial.py:102-113
 *   levels < refine_level: F.adaptive_max_pool2d; others: nearest interpolate;
 *   out = ((((0+f0)+f1)+...)+f_{L-1}) / L
 * out: device [B,C,H[refine],W[refine]] `dtype`, `layout`.
 * argmax (optional, device uint8, refine_level*B*C*Hr*Wr bytes, 4-byte aligned;
 * an opaque buffer between this call and the backward, element order
 * [level][B][C][Hr][Wr] for ARFE_NCHW and [level][B][Hr][Wr][C] for ARFE_NHWC):
 * position of the max inside the pooling window (dy*window_w+dx).  Window must
 * have <= 255 cells.
 * backward: dfeats[l] (device, `dtype`, `layout`) are fully WRITTEN. */
int arfe_fpn_gather_forward(const void* const* feats, const int32_t* H,
                            const int32_t* W, int L, int B, int C,
                            int refine_level, int dtype, int layout, void* out,
                            uint8_t* argmax, void* stream);
int arfe_fpn_gather_backward(const void* dout, const uint8_t* argmax,
                             const int32_t* H, const int32_t* W, int L, int B,
                             int C, int refine_level, int dtype, int layout,
                             void* const* dfeats, void* stream);
/* Same, with the other gradient paths of x_l folded into the one write of
 * dfeats[l]: dfeats[l] = addend[l] + (the gather's routed gradient).  x_l feeds
 * both the gather (wfpn_dual_spatial.py:102-113) and the gated residual whose
 * d x_l is d out_l itself (:135), so autograd would otherwise spend one more
 * pass over the pyramid adding the two.  addend: NULL, or L device pointers
 * (NULL entries allowed) to fp32 tensors [B,C,H[l],W[l]] in `layout`, 16-byte
 * aligned -- e.g. the gradient pyramid arfe_roi_fuse_backward_pull wrote. */
int arfe_fpn_gather_backward_acc(const void* dout, const uint8_t* argmax,
                                 const int32_t* H, const int32_t* W, int L,
                                 int B, int C, int refine_level, int dtype,
                                 int layout, const float* const* addend,
                                 void* const* dfeats, void* stream);

/* ------------------------------------------------------------------------
 * AR-FPN gated residual:
 *   out_l = x_l + nearest(bsf -> H[l] x W[l]) * (tanh(relu(g1_l)) + tanh(relu(g2_l)))
 *   WFPNDualSpatial.forward, mmdet/models/necks/wfpn_dual_spatial.py:118-135
 *   (g1_l / g2_l are the raw outputs, bias included, of the Conv2d inside
 *    reduce_convs[l] / reduce_convs2[l]; relu is mmcv ConvModule's default
 *    activation, applied here together with the tanh.)
 * feats[l], outs[l]: [B,C,H[l],W[l]]; bsf: [B,C,Hr,Wr]; g1[l], g2[l]:
 * [B,1,H[l],W[l]]; all `dtype`, feats/outs/bsf in `layout`.
 * backward (d x_l = d out_l is the identity and is left to the caller):
 *   dbsf [B,C,Hr,Wr] fp32 (written), dg1[l], dg2[l] [B,1,H[l],W[l]] fp32
 *   (written). */
int arfe_fpn_apply_forward(const void* const* feats, const void* bsf,
                           const void* const* g1, const void* const* g2,
                           const int32_t* H, const int32_t* W, int L, int B,
                           int C, int Hr, int Wr, int dtype, int layout,
                           void* const* outs, void* stream);
int arfe_fpn_apply_backward(const void* const* douts, const void* bsf,
                            const void* const* g1, const void* const* g2,
                            const int32_t* H, const int32_t* W, int L, int B,
                            int C, int Hr, int Wr, int dtype, int layout,
                            float* dbsf, float* const* dg1, float* const* dg2,
                            void* stream);

/* ------------------------------------------------------------------------
 * Proposal side: the two steps immediately before the extractor.
 *
 * arfe_nms -- greedy NMS, twin of nms_ext.nms
 * (mmdet/ops/nms/src/cuda/nms_kernel.cu:24-131, cpu/nms_cpu.cpp:8-69):
 * dets_sorted: device [n, 5] fp32 rows (x1, y1, x2, y2, score) ALREADY sorted by
 * score, descending (the reference sorts with ATen first, :79-81).  keep:
 * device int64 [n], receives the kept positions (indices into dets_sorted) in
 * order; num_keep: device int32.  Suppression rule: IoU > iou_threshold with
 * areas (x2-x1)*(y2-y1), every product and sum rounded separately as on the
 * reference's CPU path.  The bitmask is swept on the device: nothing but the
 * caller's read of num_keep crosses PCIe (the reference copies the n x n/64
 * mask to the host and sweeps it there).  workspace: arfe_nms_workspace_bytes(n)
 * bytes, 8-byte aligned.
 *
 * arfe_bbox2roi -- mmdet/core/bbox/transforms.py:41-60: rois row r of image i's
 * block = (i, x1, y1, x2, y2) of boxes[i] row r; blocks in image order.  boxes:
 * HOST array of B device pointers, counts[i] rows of cols[i] >= 4 floats each. */
size_t arfe_nms_workspace_bytes(int n);
int arfe_nms(const float* dets_sorted, int n, float iou_threshold,
             void* workspace, size_t workspace_bytes, int64_t* keep,
             int32_t* num_keep, void* stream);
int arfe_bbox2roi(const float* const* boxes, const int32_t* counts,
                  const int32_t* cols, int B, float* rois, void* stream);

/* ------------------------------------------------------------------------
 * AR-FPN gate convolutions: g1_l = reduce_convs[l].conv(x_l), g2_l =
 * reduce_convs2[l].conv(x_l) -- the two C -> 1 3x3 convolutions (padding 1, bias
 * included, no activation) of mmdet/models/necks/wfpn_dual_spatial.py:38-55,
 * :120-121, for all levels, reading every x_l ONCE (a C -> 1 convolution is a
 * channel reduction bound by reading x; through two library convolutions the
 * pyramid is read twice).  Forward only: the backward of a convolution stays
 * with the library.
 * feats[l]: [B,C,H[l],W[l]] `dtype`, channels-last (layout must be ARFE_NHWC);
 * w1[l], w2[l]: fp32 [1,C,3,3] (the Conv2d weight as stored); b1[l], b2[l]: fp32
 * [1]; g1[l], g2[l]: [B,1,H[l],W[l]] `dtype`, written.  workspace:
 * arfe_fpn_gate_conv_workspace_bytes(L, B, H, W) bytes (18 floats per pixel),
 * 16-byte aligned. */
size_t arfe_fpn_gate_conv_workspace_bytes(int L, int B, const int32_t* H,
                                          const int32_t* W);
int arfe_fpn_gate_conv_forward(const void* const* feats,
                               const float* const* w1, const float* const* b1,
                               const float* const* w2, const float* const* b2,
                               const int32_t* H, const int32_t* W, int L, int B,
                               int C, int dtype, int layout, void* workspace,
                               size_t workspace_bytes, void* const* g1,
                               void* const* g2, void* stream);

/* Backward of the same two convolutions for all levels in one pass over x:
 *   dx[l]  = sum over both filters and the 9 taps of w_f[c][tap] * dg_f[q - off(tap)]   (written;
 *            NULL dx: the inputs need no gradient)
 *   dw1[l], dw2[l] (fp32 [1,C,3,3]) and db1[l], db2[l] (fp32 [1]) are ACCUMULATED into (+=; the
 *   caller zeroes them or passes the running .grad); their summation order is not fixed.
 * feats[l], dx[l]: [B,C,H[l],W[l]] `dtype`, channels-last; dg1[l], dg2[l]: [B,1,H[l],W[l]] `dtype`
 * (gradients of the raw convolution outputs).  C a multiple of 4, at most 512. */
int arfe_fpn_gate_conv_backward(const void* const* feats,
                                const float* const* w1, const float* const* w2,
                                const void* const* dg1, const void* const* dg2,
                                const int32_t* H, const int32_t* W, int L, int B,
                                int C, int dtype, int layout, void* const* dx,
                                float* const* dw1, float* const* db1,
                                float* const* dw2, float* const* db2,
                                void* stream);

/* The whole AR-FPN backward in one pass over the incoming gradient pyramid
 * (channels-last only): what arfe_fpn_apply_backward followed by
 * arfe_fpn_gather_backward_acc(addend = douts) compute, bit for bit, reading
 * every d out element once:
 *   dbsf, dg1[l], dg2[l]  as arfe_fpn_apply_backward        (wfpn_dual_spatial.py:118-135)
 *   dx[l] = douts[l] + gradient of the gather w.r.t. x_l     (:102-113, :135)
 * douts[l]: [B,C,H[l],W[l]] in `dtype`, or fp32 when douts_f32 != 0 (the
 * accumulators arfe_roi_fuse_backward_pull wrote); dgathered: [B,C,Hr,Wr]
 * `dtype` = d(gather output); argmax from arfe_fpn_gather_forward; dx[l]:
 * `dtype`, fully written.  ARFE_E_UNSUPPORTED (nothing launched) when the case
 * is outside the fused kernel: NCHW, C % (4|8) != 0, C > 256 (fp32) / 512
 * (bf16), or a non-integer pooling ratio below the refine level. */
int arfe_fpn_backward_fused(const void* const* douts, int douts_f32,
                            const void* bsf, const void* const* g1,
                            const void* const* g2, const void* dgathered,
                            const uint8_t* argmax, const int32_t* H,
                            const int32_t* W, int L, int B, int C,
                            int refine_level, int dtype, int layout,
                            float* dbsf, float* const* dg1, float* const* dg2,
                            void* const* dx, void* stream);

/* ------------------------------------------------------------------------
 * NonLocal2D refine as one fused attention on the tensor cores (tcgen05, TMEM):
 *   y[b, p, :] = sum_q softmax_q(scale * <theta[b, p, :], phi[b, q, :]>) g[b, q, :]
 * = mmdet/ops/non_local.py:65-69 (embedded_gaussian; scale = 1, or
 * 1/sqrt(inter_channels) when use_scale) followed by :98-101 (pairwise_weight .
 * g_x), the call of mmdet/models/necks/wfpn_dual_spatial.py:115.  The HW x HW
 * weight matrix (70.6 MB per image at 50 x 84) is never materialised.  The 1x1
 * convolutions g / theta / phi / conv_out and the residual stay with the caller.
 * theta, phi, g, y: [B, D, H, W] `dtype` with HW = H * W, all ARFE_NCHW or all
 * ARFE_NHWC (y has the layout of the inputs and is what the reference's
 * `y.permute(0, 2, 1).contiguous().reshape(n, D, h, w)` holds); D = inter_channels
 * in {64, 128, 256}.  Operands are rounded to bf16, accumulation and softmax are
 * fp32: results agree with the fp32 reference to bf16 accuracy (1e-2).  All
 * four tensors 16-byte aligned.  theta is always read in place; bf16 ARFE_NHWC
 * phi / g are read in place too (tensor maps), any other combination of phi / g
 * is first converted into bf16 tiles in the workspace.
 * nsplit >= 1 slices the key range over several CTAs per 128-query block
 * (arfe_nonlocal_default_split: enough to fill the SMs of the current device);
 * workspace: arfe_nonlocal_workspace_bytes(B, HW, D, nsplit) bytes, 1024-byte
 * aligned (bf16 operand tiles + the partial results and arrival counters when
 * nsplit > 1); it holds all device-side state of a call: concurrent calls on
 * different streams need different workspaces.  scale must be positive.
 * Forward only. */
int arfe_nonlocal_default_split(int B, int HW);
size_t arfe_nonlocal_workspace_bytes(int B, int HW, int D, int nsplit);
int arfe_nonlocal_attention_forward(const void* theta, const void* phi,
                                    const void* g, void* y, int B, int HW, int D,
                                    int dtype, int layout, float scale,
                                    int nsplit, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Row pass of the attention backward (the five GEMMs around it are library calls on bf16 operands): for
 * every query row, from the logits S = theta_x . phi_x and dP = dY . g_x^T (fp32 [rows][n], read once),
 *   P = softmax(scale * S),  dS = scale * P * (dP - sum_q P dP)     (non_local.py:65-69 differentiated)
 * written as bf16 [rows][n].  n <= 51200. */
int arfe_nonlocal_backward_rows(const float* S, const float* dP, void* P_bf16,
                                void* dS_bf16, int64_t rows, int n, float scale,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARFE_B200_H_ */
