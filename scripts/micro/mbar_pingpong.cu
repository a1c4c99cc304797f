// Micro-benchmark: round-trip latency of a producer/consumer hand-off inside one CTA
//   (a) mbarrier.arrive -> mbarrier.try_wait (both directions)
//   (b) shared-memory counters: atomicAdd / volatile polling
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_pingpong mbar_pingpong.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t par) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
  } while (!done);
}
// warp 0 = producer, warps 1..NC = consumers (each arrives once per round)
template <int MODE>
__global__ void pingpong(int rounds, int ncons, long long* out) {
  __shared__ uint64_t full, empty;
  __shared__ volatile int f_cnt, e_cnt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mb_init(&full, 1); mb_init(&empty, ncons); f_cnt = 0; e_cnt = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long t0 = clock64();
  if (MODE == 0) {
    if (warp == 0) {
      for (int i = 0; i < rounds; ++i) {
        if (lane == 0) mb_arrive(&full);
        mb_wait(&empty, i & 1);
      }
    } else if (warp <= ncons) {
      for (int i = 0; i < rounds; ++i) {
        mb_wait(&full, i & 1);
        __syncwarp();
        if (lane == 0) mb_arrive(&empty);
      }
    }
  } else {
    if (warp == 0) {
      for (int i = 0; i < rounds; ++i) {
        if (lane == 0) { __threadfence_block(); f_cnt = i + 1; }
        while (e_cnt < (i + 1) * ncons) {}
        __syncwarp();
      }
    } else if (warp <= ncons) {
      for (int i = 0; i < rounds; ++i) {
        while (f_cnt < i + 1) {}
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd((int*)&e_cnt, 1); }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 1024);
  for (int ncons : {1, 7, 8, 14}) {
    for (int mode = 0; mode < 2; ++mode) {
      const int rounds = 20000;
      for (int blocks : {1, 296}) {
        if (mode == 0) pingpong<0><<<blocks, 32 * (ncons + 1)>>>(rounds, ncons, d);
        else pingpong<1><<<blocks, 32 * (ncons + 1)>>>(rounds, ncons, d);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("consumers %2d  %s  blocks %3d : %6.0f cycles per round trip (%s)\n", ncons, mode == 0 ? "mbarrier" : "smem flag", blocks,
               (double)h / rounds, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
