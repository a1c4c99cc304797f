// Micro-benchmark: how fast can cp.async.bulk (TMA, 1-D) move L2-resident data into
// shared memory on this GPU?  This is the ceiling of the ring forward / pull backward,
// whose windows overlap ~6.5x and are therefore served from L2, not HBM.
//   one elected thread per CTA keeps `slots` copies of `piece` bytes in flight
//   (mbarrier per slot), nothing consumes the data.
//   pattern 0: CTAs stream disjoint pieces of a buffer that fits L2 (resident after warm-up)
//   pattern 1: "window rows": each CTA copies runs of `rows` pieces `row_pitch` apart
//              from a random start (the access shape of the ring forward)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_tma_bw l2_tma_bw.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t par) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

constexpr int kMaxBars = 1024;

// blockDim = warps * 32; issuers = lanes < `lanes` of every warp, each with its own slots.
__global__ void __launch_bounds__(256) stream_kernel(const char* buf, size_t bytes, int piece, int slots, int iters,
                                                     int pattern, int rows, size_t row_pitch, int lanes, int mode) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bars[kMaxBars];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int me = warp * lanes + lane;          // issuer id
  const int issuers = nwarps * lanes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < issuers * slots; ++s) mb_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane >= lanes) return;
  uint64_t* bar = bars + me * slots;
  char* ring = smem + (size_t)me * slots * piece;
  const size_t npieces = bytes / piece;
  const size_t stride = (size_t)gridDim.x * issuers;
  size_t idx = (size_t)blockIdx.x * issuers + me;
  uint32_t rng = 747796405u * (uint32_t)(idx + 1) + 2891336453u;
  int row = 0;
  size_t base = 0;
  auto next = [&]() -> const char* {
    if (pattern == 0) {
      const char* p = buf + (idx % npieces) * (size_t)piece;
      idx += stride;
      return p;
    }
    if (row == 0) {
      rng = rng * 1664525u + 1013904223u;
      const size_t span = bytes - (size_t)rows * row_pitch - piece;
      base = ((size_t)(rng >> 4) * 1024) % span;
      base &= ~(size_t)15;
    }
    const char* p = buf + base + (size_t)row * row_pitch;
    if (++row == rows) row = 0;
    return p;
  };
  for (int s = 0; s < slots; ++s) {
    mb_expect(&bar[s], piece);
    bulk(ring + (size_t)s * piece, next(), piece, &bar[s]);
  }
  int s = 0;
  uint32_t ph = 0;
  if (mode == 1) {        // barrier operations alone: wait + arrive, no copy
    for (int it = 0; it < iters; ++it) {
      mb_wait(&bar[s], ph);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[s])) : "memory");
      if (++s == slots) { s = 0; ph ^= 1; }
    }
    return;
  }
  for (int it = 0; it < iters; ++it) {
    mb_wait(&bar[s], ph);
    mb_expect(&bar[s], piece);
    if (mode == 2) {
      const char* src = next();
      bulk(ring + (size_t)s * piece, src, piece / 2, &bar[s]);
      bulk(ring + (size_t)s * piece + piece / 2, src + piece / 2, piece / 2, &bar[s]);
    } else {
      bulk(ring + (size_t)s * piece, next(), piece, &bar[s]);
    }
    if (++s == slots) { s = 0; ph ^= 1; }
  }
  for (int k = 0; k < slots; ++k) {
    mb_wait(&bar[s], ph);
    if (++s == slots) { s = 0; ph ^= 1; }
  }
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int pieces[] = {512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
  const int shapes[][2] = {{1, 1}, {2, 1}, {4, 1}, {8, 1}, {1, 4}, {1, 32}, {4, 8}};  // warps, lanes
  char* buf = nullptr;
  cudaMalloc(&buf, (size_t)512 << 20);
  cudaMemset(buf, 1, (size_t)512 << 20);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("mode,pattern,buffer_MB,piece_B,ring_KB,ctas_per_sm,warps,lanes,slots_per_issuer,GBps,ns_per_copy_per_cta\n");
  const int nmodes = argc > 1 ? atoi(argv[1]) : 1;  // 0 stream, 1 barrier ops only, 2 two bulk copies per slot
  for (int mode = 0; mode < nmodes; ++mode)
  for (int pattern = 0; pattern < 2; ++pattern)
    for (size_t mb : {(size_t)64, (size_t)400})
      for (int piece : pieces)
        for (auto& sh : shapes)
          for (int cps : {1, 2}) {
            if (mode > 0 && (pattern != 0 || mb != 64 || piece > 4096)) continue;
            const int warps = sh[0], lanes = sh[1], issuers = warps * lanes;
            const int ring = (cps == 1 ? 192 : 96) * 1024;
            int slots = ring / (piece * issuers);
            if (slots < 2) continue;
            if (slots * issuers > kMaxBars) slots = kMaxBars / issuers;
            if (slots > 64) slots = 64;
            const size_t bytes = mb << 20;
            const int grid = sms * cps;
            long long iters = ((long long)2 << 30) / ((long long)grid * issuers * piece);
            if (iters < 8) iters = 8;
            const size_t smem = (size_t)slots * issuers * piece;
            const int rows = 20;
            const size_t pitch = 84 * 1024;  // one row of the stride-16 level at C = 256 fp32
            stream_kernel<<<grid, warps * 32, smem>>>(buf, bytes, piece, slots, (int)(iters / 8 + 1), pattern, rows, pitch, lanes, mode);
            cudaEventRecord(e0);
            stream_kernel<<<grid, warps * 32, smem>>>(buf, bytes, piece, slots, (int)iters, pattern, rows, pitch, lanes, mode);
            cudaEventRecord(e1);
            cudaError_t err = cudaEventSynchronize(e1);
            if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double copies = (double)issuers * (iters + slots);
            const double moved = (double)grid * copies * piece;
            printf("%d,%d,%zu,%d,%d,%d,%d,%d,%d,%.0f,%.1f\n", mode, pattern, mb, piece, ring / 1024, cps, warps, lanes, slots,
                   moved / ms * 1e-6, ms * 1e6 / copies);
            fflush(stdout);
          }
  return 0;
}
