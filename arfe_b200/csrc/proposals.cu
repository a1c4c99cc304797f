// Proposal side of the path: the two steps that sit
// immediately before the fused extractor.
//
//   nms      : mmdet/ops/nms/src/cuda/nms_kernel.cu:24-67 builds a 64 x 64-tiled
//              suppression bitmask on the device, copies it to the HOST
//              (boxes * ceil(boxes / 64) * 8 bytes: 3 MB at 5000 boxes) and sweeps it
//              there (:96-123).  Here the sweep stays on the device: only the
//              upper-triangular tiles are computed (the sweep never reads the others),
//              one CTA walks the tiles in order -- a 64-step register sweep over the
//              diagonal tile decides which boxes of the tile survive, then all threads
//              OR the survivors' mask rows into the removed set in parallel -- and what
//              crosses PCIe is the count.  IoU with the reference's arithmetic
//              (no FMA contraction: areas and the intersection are rounded products, as
//              in nms_cpu.cpp:21,59-62) so that "> threshold" agrees with it.
//   bbox2roi : mmdet/core/bbox/transforms.py:41-60 -- per image new_full + cat, then a
//              cat over images -- as one launch writing the [n, 5] RoI tensor in the
//              image-major order the extractor's plan relies on.
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.h"

namespace arfe {
namespace {

constexpr int kTile = 64;  // boxes per mask word

__device__ __forceinline__ float iou_ref(const float* a, const float* b) {
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
  const float sb = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

// grid (col tiles, row tiles), 64 threads; only tiles on or above the diagonal
__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float* __restrict__ dets, int n, float thr, unsigned long long* __restrict__ mask) {
  const int row_t = blockIdx.y, col_t = blockIdx.x;
  if (row_t > col_t) return;
  const int row_n = min(n - row_t * kTile, kTile), col_n = min(n - col_t * kTile, kTile);
  __shared__ float cb[kTile * 5];
  if ((int)threadIdx.x < col_n) {
#pragma unroll
    for (int q = 0; q < 5; ++q) cb[threadIdx.x * 5 + q] = dets[(size_t)(col_t * kTile + threadIdx.x) * 5 + q];
  }
  __syncthreads();
  if ((int)threadIdx.x < row_n) {
    const int i = row_t * kTile + threadIdx.x;
    float me[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) me[q] = dets[(size_t)i * 5 + q];
    unsigned long long t = 0;
    const int start = row_t == col_t ? threadIdx.x + 1 : 0;
    for (int j = start; j < col_n; ++j)
      if (iou_ref(me, cb + j * 5) > thr) t |= 1ull << j;
    const int cols = (n + kTile - 1) / kTile;
    mask[(size_t)i * cols + col_t] = t;
  }
}

// One CTA.  removed[] lives in shared memory (one word per tile).
constexpr int kSweepThreads = 256;
__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int n, int64_t* __restrict__ keep,
                 int* __restrict__ num_keep) {
  extern __shared__ unsigned long long removed[];  // [cols]
  __shared__ unsigned long long diag[kTile];
  __shared__ unsigned long long alive_s;
  __shared__ int base_s;
  const int cols = (n + kTile - 1) / kTile;
  const int tid = threadIdx.x;
  for (int j = tid; j < cols; j += kSweepThreads) removed[j] = 0ull;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int t = 0; t < cols; ++t) {
    const int cnt = min(n - t * kTile, kTile);
    if (tid < cnt) diag[tid] = mask[(size_t)(t * kTile + tid) * cols + t];
    __syncthreads();
    if (tid == 0) {
      // boxes of this tile in score order: a box survives unless an earlier survivor removed it
      unsigned long long rem = removed[t], alive = 0ull;
      for (int i = 0; i < cnt; ++i)
        if (!((rem >> i) & 1ull)) { alive |= 1ull << i; rem |= diag[i]; }
      alive_s = alive;
    }
    __syncthreads();
    const unsigned long long alive = alive_s;
    const int base = base_s;
    // survivors -> keep list (position in the score order), in order
    if (tid < cnt && ((alive >> tid) & 1ull))
      keep[base + __popcll(alive & ((1ull << tid) - 1ull))] = (int64_t)t * kTile + tid;
    // their mask rows -> removed set of the tiles to the right
    for (int j = t + 1 + tid; j < cols; j += kSweepThreads) {
      unsigned long long acc = removed[j];
      unsigned long long a = alive;
      while (a) {
        const int i = __ffsll((long long)a) - 1;
        a &= a - 1ull;
        acc |= mask[(size_t)(t * kTile + i) * cols + j];
      }
      removed[j] = acc;
    }
    __syncthreads();
    if (tid == 0) base_s = base + __popcll(alive);
    __syncthreads();
  }
  if (tid == 0) *num_keep = base_s;
}

constexpr int kMaxLists = 64;
struct BoxLists {
  const float* ptr[kMaxLists];
  int start[kMaxLists + 1];  // first output row of list i
  int cols[kMaxLists];       // floats per input row (>= 4)
  int img0;                  // image id of list 0
  int n;
};

__global__ void __launch_bounds__(256)
bbox2roi_kernel(const BoxLists bl, float* __restrict__ rois) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= bl.start[bl.n]) return;
  int i = 0;
  while (i + 1 < bl.n && r >= bl.start[i + 1]) ++i;
  const float* b = bl.ptr[i] + (size_t)(r - bl.start[i]) * bl.cols[i];
  float* o = rois + (size_t)r * 5;
  o[0] = (float)(bl.img0 + i);
  o[1] = b[0]; o[2] = b[1]; o[3] = b[2]; o[4] = b[3];
}

}  // namespace

size_t nms_workspace_bytes(int n) {
  const size_t cols = (size_t)(n + kTile - 1) / kTile;
  return (size_t)n * cols * sizeof(unsigned long long);
}

cudaError_t launch_nms(const float* dets_sorted, int n, float thr, void* workspace, int64_t* keep, int* num_keep,
                       cudaStream_t stream) {
  if (n == 0) return cudaMemsetAsync(num_keep, 0, sizeof(int), stream);
  const int cols = (n + kTile - 1) / kTile;
  if ((size_t)cols * 8 > 160 * 1024) return cudaErrorInvalidValue;  // > 1.3 M boxes
  auto* mask = static_cast<unsigned long long*>(workspace);
  nms_mask_kernel<<<dim3(cols, cols), kTile, 0, stream>>>(dets_sorted, n, thr, mask);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int smem = cols * 8;
  if (smem > 48 * 1024 &&
      (e = cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
    return e;
  nms_sweep_kernel<<<1, kSweepThreads, smem, stream>>>(mask, n, keep, num_keep);
  return cudaGetLastError();
}

cudaError_t launch_bbox2roi(const float* const* boxes, const int* counts, const int* cols, int B, float* rois,
                            cudaStream_t stream) {
  int done = 0, row0 = 0;
  while (done < B) {
    BoxLists bl;
    const int nb = B - done < kMaxLists ? B - done : kMaxLists;
    bl.n = nb;
    bl.img0 = done;
    int tot = 0;
    for (int i = 0; i < kMaxLists; ++i) {
      bl.ptr[i] = i < nb ? boxes[done + i] : nullptr;
      bl.cols[i] = i < nb ? cols[done + i] : 4;
      bl.start[i] = tot;
      if (i < nb) tot += counts[done + i];
    }
    bl.start[kMaxLists] = tot;
    for (int i = nb; i <= kMaxLists; ++i) bl.start[i] = tot;
    if (tot > 0) {
      bbox2roi_kernel<<<(tot + 255) / 256, 256, 0, stream>>>(bl, rois + (size_t)row0 * 5);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
    row0 += tot;
    done += nb;
  }
  return cudaSuccess;
}

}  // namespace arfe
