// Hand-off latencies inside one CTA on sm_100a (what paces a step of the attention kernel):
//   (a) mbarrier.arrive by warp 1            -> warp 0 (try_wait) sees it
//   (b) tcgen05.commit (no MMA in flight)    -> warp 0 sees the mbarrier
//   (c) one tcgen05.mma M128 N64 K16 + commit -> warp 0 sees it   (pipeline latency of a single MMA)
//   (d) 16 dependent MMAs (one accumulator) + commit, (e) 4 MMAs M128 N256 K16 + commit
// clock64 of the same SM on both sides.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tc_commit_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(par) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) { return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(64, 1) k(long long* out, int mode, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];  // 64 KB of (zero) operands
  __shared__ uint64_t go, done;
  __shared__ uint32_t slot;
  __shared__ long long t_issue;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&go, 1); mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  long long acc = 0;
  for (int r = 0; r < reps; ++r) {
    const uint32_t par = r & 1;
    if (warp == 0) {  // waiter: releases warp 1, then waits for its signal
      if (lane == 0) mbar_arrive(&go);
      mbar_wait(&done, par);
      const long long t1 = clock64();
      if (lane == 0) acc += t1 - *reinterpret_cast<volatile long long*>(&t_issue);
      __syncwarp();
    } else {          // signaller
      mbar_wait(&go, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a = smem_u32(smem);
        *reinterpret_cast<volatile long long*>(&t_issue) = clock64();
        if (mode == 0) mbar_arrive(&done);
        else {
          if (mode == 2) mma_ss(tm + 256, desc_sw128(a), desc_sw128(a + 16384), idesc(128, 64), 0);
          if (mode == 3) for (int kk = 0; kk < 16; ++kk) mma_ss(tm + 256, desc_sw128(a + (kk & 3) * 32), desc_sw128(a + 16384 + (kk & 3) * 32), idesc(128, 64), kk > 0);
          if (mode == 4) for (int kk = 0; kk < 4; ++kk) mma_ss(tm, desc_sw128(a + kk * 32), desc_sw128(a + 32768 + kk * 32), idesc(128, 256), kk > 0);
          tc_commit(&done);
        }
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0) out[blockIdx.x] = acc / reps;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* names[] = {"mbarrier.arrive -> waiter", "tcgen05.commit (empty pipe) -> waiter", "1 MMA M128 N64 K16 + commit -> waiter",
                         "16 dependent MMAs M128 N64 K16 + commit -> waiter", "4 dependent MMAs M128 N256 K16 + commit -> waiter"};
  for (int grid : {1, 132})
    for (int mode = 0; mode < 5; ++mode) {
      k<<<grid, 64, 65536>>>(d, mode, 200);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0, sum = 0;
      for (int i = 0; i < grid; ++i) { sum += h[i]; mx = h[i] > mx ? h[i] : mx; }
      printf("grid %3d  %-52s %6lld cycles (max over CTAs %lld)%s\n", grid, names[mode], sum / grid, mx, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
