// NonLocal2D refine of the AR-FPN neck as ONE fused attention on the 5th-generation tensor
// cores.
//
// Reference: mmdet/ops/non_local.py:65-69 (embedded_gaussian: softmax(theta_x . phi_x [/ sqrt(Ci)]))
// and :78-101 (y = pairwise_weight . g_x), called from necks/wfpn_dual_spatial.py:115 with
// reduction = 1, use_scale = False: per image a 4200 x 4200 x 256 attention whose weight matrix the
// reference materialises in fp32 (70.6 MB per image, written and re-read three times).
//
// What runs here (the 1x1 convolutions theta / phi / g / conv_out stay plain library GEMMs):
//
//  nl_pack_kernel      phi, g (fp32 or bf16, NCHW or channels-last) -> bf16 tiles of 64 positions that are
//                      already the shared-memory image the tensor core wants (128-byte swizzle: 16-byte
//                      chunk c of row r stored at chunk c ^ (r & 7)), one contiguous blob per tile, so the
//                      attention kernel fetches a K or V tile with a single bulk copy (cp.async.bulk)
//                      straight into place.  Tiles keep the orientation of the input -- the MMA takes a
//                      B operand K-major or MN-major -- so nothing is transposed:
//                        channels-last: [B][nkb][D/64 slabs][64 position rows][128 B = 64 channels]
//                                       (phi: K-major operand of Q K^T, g: MN-major operand of P V)
//                        NCHW:          [B][nkb][D channel rows][128 B = 64 positions]
//                                       (phi: MN-major operand of Q K^T, g: K-major operand of P V)
//                      bf16 channels-last inputs skip this pass: their tiles are read in place through
//                      tensor maps (cp.async.bulk.tensor, hardware swizzle, zero fill past HW).
//                      theta is never packed: the attention kernel reads it in place (any layout / type),
//                      converts and writes it to tensor memory.
//  nl_attn_kernel      one CTA = 128 query positions of one image (x one slice of the keys when the
//                      key range is split to fill the SMs).  Warp roles:
//                        warps 0-7  softmax, thread = (query row, half of the 64 key columns): S (TMEM) ->
//                                   registers, running max / sum, P (bf16) -> TMEM over S; rescale of O in TMEM
//                                   only when a row maximum grew by more than 2^8 (exact: the maximum
//                                   used as the exponent's reference is arbitrary); epilogue O / l -> y
//                        warp 8     one lane issues tcgen05.mma: S = Q K^T (M128 N64 K16 x D/16) and
//                                   O += P V (M128 N=D K16 x 4); both A operands (Q, P) come from tensor
//                                   memory, accumulators in tensor memory; tcgen05.commit -> mbarriers
//                        warps 9,10 one lane each issues the bulk copies of the K ring / the V ring
//                      Q is written to TMEM once (as bf16 pairs) by the softmax warps: with A in shared
//                      memory every M128 N64 K16 instruction fetched 4 KB of Q for 32 cycles of math and
//                      the tensor pipe ran at half speed.  Two S tiles live in TMEM (Q K^T of step j + 1
//                      runs under the softmax of step j); P overwrites its own S tile, so the softmax
//                      weights never touch shared memory (four S tiles, Q K^T three steps ahead, where
//                      tensor memory has room: D <= 128); K and V rings are three tiles deep.  One
//                      tcgen05.commit per MMA group: completion barriers in rings of six, shared by
//                      the softmax warps and the tile loaders.
//                      Key range split: each CTA leaves (O, max, sum) in the workspace and the CTA of a
//                      query block that arrives last (an atomic counter; nobody waits) merges them and
//                      writes y -- no second kernel.
//
// Arithmetic: operands rounded to bf16 (round to nearest even), products exact, fp32 accumulation in
// the tensor core, fp32 softmax with exp2; tolerance against the fp32 reference is the bf16 one of
// north_star (1e-2), written in tests/test_nonlocal_gpu.py.
#include <cuda.h>
#include <cstdio>
#include <string.h>
#include <cuda_bf16.h>

#include "launch.h"
#include "tma.cuh"

namespace arfe {
namespace {

constexpr int NL_BM = 128;  // queries per CTA = UMMA M
constexpr int NL_BN = 64;   // keys per step = one 128-byte swizzle row of bf16
constexpr float NL_RESCALE = 8.f;  // log2 units a row maximum may grow before O is rescaled
constexpr int NL_SBUF_MAX = 4;     // S tiles in tensor memory (as many as fit): Q K^T runs SBUF - 1 steps ahead of P V
constexpr int NL_STAGES = 3;       // K and V tile rings in shared memory
// phi (K) and g (V) tiles share one layout per input layout (rows = positions for channels-last inputs,
// rows = channels for NCHW ones); the MMA takes one of them as a K-major and the other as an MN-major B
// operand, so nothing is ever transposed.

template <int D>
struct NlCfg {
  static constexpr int SLABS = D / 64;
  static constexpr int K_BYTES = NL_BN * D * 2;
  static constexpr int V_BYTES = D * NL_BN * 2;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + NL_STAGES * K_BYTES;
  static constexpr int OFF_BAR = OFF_V + NL_STAGES * V_BYTES;
  static constexpr int OFF_XCH = OFF_BAR + 256;      // row max / sum exchange between the two column halves
  static constexpr int SMEM = OFF_XCH + 2 * 2 * NL_BM * 4;  // the kernel has no static shared memory: base 1024-aligned
  // tensor memory columns: O (D, fp32) | Q (D / 2: bf16 pairs) | S0 .. (64 each; P_j is written over S_j)
  static constexpr int TM_Q = D, TM_S = D + D / 2;
  static constexpr int SBUF = (512 - TM_S) / NL_BN < NL_SBUF_MAX ? (512 - TM_S) / NL_BN : NL_SBUF_MAX;  // 2 at D = 256
  static constexpr int TMEM_USED = TM_S + SBUF * NL_BN;
  static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512);
};

// ---- tcgen05 / TMEM primitives (inline PTX) ---------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a CONVERGED warp.  The single-thread tcgen05 instructions must sit in warp-uniform
// control flow behind this predicate: inside a divergent `lane == 0` branch the compiler wraps every
// tcgen05.mma / commit in an elect-and-retry loop with register -> uniform-register moves (about 70
// cycles per instruction: the issuing thread, not the tensor pipe, then paces the kernel).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {  // arrives when all MMAs issued so far have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 operands, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with A read from tensor memory (row m in lane m, elements 2 c and 2 c + 1 of the row in
// the low and high half of 32-bit column c): D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 | SBO >> 4 at bit 32 | version 1 at bit 46 | SWIZZLE_128B = 2 at bit 61)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same for an MN-major operand (the N / M index contiguous): 8 k-rows of 128 bytes form a swizzle
// atom of 64 MN-elements; LBO = bytes between atoms along MN, SBO = bytes between 8-row groups along K
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulator, bf16 A and B; A K-major,
// B K-major or MN-major (bit 16)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

#define NL_R32(r)                                                                                                  \
  r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16],    \
      r[17], r[18], r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]
// 32 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32 (w % 4) ...)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// A wait that cannot hang the device: a protocol error traps (the launch fails with an error the
// C ABI reports) instead of spinning for ever.
__device__ __forceinline__ void nl_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && (spin & 1023u) == 1023u) {  // 20 s without progress: a protocol error, not a slow step
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 20000000000ull) __trap();
    }
  }
}

// 3-D tensor-map copy (TMA proper): box (64 channels, 64 positions, 1 image) of a [B][HW][D] bf16 tensor,
// written with the 128-byte swizzle the MMA descriptors expect; positions beyond HW are zero-filled.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Phase-skip knob of the -DARFE_PROFILE build (scripts/nonlocal_knobs.py), bits: 1 no K / V copies after
// the ring's first fill, 2 no softmax work at all, 4 no MMAs, 16 no row-max exchange between the column
// halves, 32 no ex2, 64 no P store, 128 no S load, 1024 two key steps only (fixed cost of a CTA).
// Results are garbage (without MMAs S is stale: data-dependent paths like the O rescale then run at
// random, so those timings say little).  Folds to false in the shipped library.
#ifdef ARFE_PROFILE
#define NL_DBG(bit) ((dbg & (bit)) != 0)
#else
#define NL_DBG(bit) false
#endif

// Per-phase clock accounting of the softmax loop (profile build, ARFE_NL_DBG bit 65536): block (0,0,0),
// warp 0, lane 0 prints the average cycles per step of each phase.
#ifdef ARFE_PROFILE
#define NL_T(i) do { if (NL_DBG(65536)) { const long long t_ = clock64(); tacc[i] += t_ - tlast; tlast = t_; } } while (0)
#else
#define NL_T(i) do { } while (0)
#endif

// barrier indices
// One tcgen05.commit per MMA group (a commit costs the issuing warp several hundred cycles: four per step
// paced the whole kernel): "Q K^T of step it done" (B_QKDONE) tells the softmax that S is ready AND the K
// loader that the tile's slot is free; "P V of step it done" (B_PVDONE) frees the V slot and O.  Both are
// rings of NL_EVT = 6 barriers (a multiple of the 2 S tiles and the 3 ring slots): completion it + 6 on a
// barrier needs tile it + 6, which its loader fetches only after it has seen completion it + 3 > it -- no
// waiter can be lapped, so the parity waits stay exact.
constexpr int NL_EVT = 6;
enum { B_QFULL = 0, B_KFULL = 1, B_VFULL = 4, B_PFULL = 7, B_QKDONE = 11, B_PVDONE = 17, B_COUNT = 23 };

template <typename OutT>
__device__ __forceinline__ void nl_store1(OutT* p, float v);
template <>
__device__ __forceinline__ void nl_store1<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void nl_store1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename T>
__device__ __forceinline__ float nl_to_float(T v);
template <>
__device__ __forceinline__ float nl_to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float nl_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------
constexpr int NL_SOFTMAX_WARPS = 8;  // two per TMEM lane quadrant: each takes half of the 64 key columns
constexpr int NL_THREADS = (NL_SOFTMAX_WARPS + 3) * 32;  // + MMA issuer, K loader, V loader

// TM = false: Qp / Kp / Vp are the packed operands of nl_pack_kernel (tiles by 1-D bulk copies).
// TM = true : the inputs are bf16 channels-last already ([B][HW][D] rows): Qp is theta itself, and the K / V
//             tiles come straight from phi / g through tensor maps (cp.async.bulk.tensor, hardware swizzle,
//             zero fill past HW) -- no packing pass at all.
template <int D, typename OutT, bool TM, bool POS>
__global__ void __launch_bounds__(NL_THREADS, 1)
nl_attn_kernel(const uint8_t* __restrict__ Qp, const uint8_t* __restrict__ Kp, const uint8_t* __restrict__ Vp,
               const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
               OutT* __restrict__ y, float* __restrict__ part_o, float* __restrict__ part_ml,
               int* __restrict__ counters, int HW, int nqb, int nkb, int nsplit, float sl2, int out_cl, int dbg) {
  using C = NlCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128-byte swizzle atoms need a 1024-byte aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* xch = reinterpret_cast<float*>(smem + C::OFF_XCH);  // [2 parities][2 halves][128 rows]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x, b = blockIdx.y, z = blockIdx.z, B = gridDim.y;
  const int kb_lo = (int)((long long)nkb * z / nsplit), kb_hi = (int)((long long)nkb * (z + 1) / nsplit);
  const int n_it = (NL_DBG(1024) && kb_hi - kb_lo > 2) ? 2 : kb_hi - kb_lo;  // 1024: fixed cost only

  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i)
      mbar_init(&bars[i], ((i >= B_PFULL && i < B_PFULL + NL_SBUF_MAX) || i == B_QFULL) ? NL_SOFTMAX_WARPS : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NL_SOFTMAX_WARPS) {  // tensor memory: O | Q | S0 | S1
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t tmem_o = tmem, tmem_q = tmem + C::TM_Q, tmem_s = tmem + C::TM_S;

  if (warp == NL_SOFTMAX_WARPS + 1) {
    // ===== loader of the K ring (a slot is free once the Q K^T that read it has completed) =====
    for (int it = 0; it < n_it; ++it) {
      const int s = it % NL_STAGES;
      if (it >= NL_STAGES)  // the slot's previous tile: Q K^T of step it - NL_STAGES done
        nl_wait(&bars[B_QKDONE + (it - NL_STAGES) % NL_EVT], (uint32_t)((it - NL_STAGES) / NL_EVT) & 1u);
      if (elect_one()) {
        if (NL_DBG(1) && it >= NL_STAGES) {
          mbar_arrive(&bars[B_KFULL + s]);
        } else {
          mbar_arrive_expect_tx(&bars[B_KFULL + s], C::K_BYTES);
          if constexpr (TM) {
#pragma unroll
            for (int sl = 0; sl < C::SLABS; ++sl)
              tma_load_3d(smem + C::OFF_K + s * C::K_BYTES + sl * (NL_BN * 128), &tm_k, sl * 64, (kb_lo + it) * NL_BN, b,
                          &bars[B_KFULL + s]);
          } else {
            bulk_g2s(smem + C::OFF_K + s * C::K_BYTES, Kp + ((size_t)b * nkb + (kb_lo + it)) * C::K_BYTES, C::K_BYTES,
                     &bars[B_KFULL + s]);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == NL_SOFTMAX_WARPS + 2) {
    // ===== loader of the V ring (free once the P V that read it has completed) =====
    for (int it = 0; it < n_it; ++it) {
      const int s = it % NL_STAGES;
      if (it >= NL_STAGES)  // P V of step it - NL_STAGES done
        nl_wait(&bars[B_PVDONE + (it - NL_STAGES) % NL_EVT], (uint32_t)((it - NL_STAGES) / NL_EVT) & 1u);
      if (elect_one()) {
        if (NL_DBG(1) && it >= NL_STAGES) {
          mbar_arrive(&bars[B_VFULL + s]);
        } else {
          mbar_arrive_expect_tx(&bars[B_VFULL + s], C::V_BYTES);
          if constexpr (TM) {
#pragma unroll
            for (int sl = 0; sl < C::SLABS; ++sl)
              tma_load_3d(smem + C::OFF_V + s * C::V_BYTES + sl * (NL_BN * 128), &tm_v, sl * 64, (kb_lo + it) * NL_BN, b,
                          &bars[B_VFULL + s]);
          } else {
            bulk_g2s(smem + C::OFF_V + s * C::V_BYTES, Vp + ((size_t)b * nkb + (kb_lo + it)) * C::V_BYTES, C::V_BYTES,
                     &bars[B_VFULL + s]);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == NL_SOFTMAX_WARPS) {
    // ===== MMA issuer: the whole warp walks the schedule (warp-uniform), one elected lane issues =====
    // tiles with position rows (channels-last inputs): K is the K-major, V the MN-major B operand;
    // tiles with channel rows (NCHW inputs): the other way round
    constexpr bool pos_rows = POS;  // == (out_cl != 0): the layout of the inputs
    constexpr uint32_t idesc_qk = idesc_bf16(NL_BM, NL_BN, !pos_rows);
    constexpr uint32_t idesc_pv = idesc_bf16(NL_BM, D, pos_rows);
    nl_wait(&bars[B_QFULL], 0);  // Q sits in tensor memory (written by the softmax warps)
    tc_fence_after();
    // S[it % NL_SBUF] = Q K_it^T: A = Q from tensor memory (no shared-memory operand traffic for the
    // 128 x 16 A slice of every instruction), B = K tile
    auto issue_qk = [&](int it) {
      const int s = it % NL_STAGES;
      nl_wait(&bars[B_KFULL + s], (uint32_t)(it / NL_STAGES) & 1u);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(smem + C::OFF_K + s * C::K_BYTES);
      const uint32_t d_tmem = tmem_s + (uint32_t)(it % C::SBUF) * NL_BN;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t bd = pos_rows ? smem_desc_sw128(k_addr + (kk >> 2) * (NL_BN * 128) + (kk & 3) * 32)
                                       : smem_desc_sw128_mn(k_addr + kk * (16 * 128), 1024, 1024);  // 16 channel rows on
          if (!NL_DBG(4)) tc_mma_ts(d_tmem, tmem_q + kk * 8, bd, idesc_qk, kk > 0);
        }
        tc_commit(&bars[B_QKDONE + it % NL_EVT]);
      }
      __syncwarp();
    };
    for (int it = 0; it < C::SBUF - 1 && it < n_it; ++it) issue_qk(it);
    for (int jt = 0; jt < n_it; ++jt) {
      // the Q K^T SBUF - 1 steps ahead goes first: its S tile held P_{jt-1}, which the P V issued
      // in the previous round has consumed (the tensor pipe executes in issue order)
      if (jt + C::SBUF - 1 < n_it) issue_qk(jt + C::SBUF - 1);
      // O += P_jt V_jt
      const int s = jt % NL_STAGES;
      nl_wait(&bars[B_VFULL + s], (uint32_t)(jt / NL_STAGES) & 1u);
      nl_wait(&bars[B_PFULL + jt % C::SBUF], (uint32_t)(jt / C::SBUF) & 1u);
      tc_fence_after();
      const uint32_t p_tmem = tmem_s + (uint32_t)(jt % C::SBUF) * NL_BN;  // P_jt lies over S_jt
      const uint32_t v_addr = smem_u32(smem + C::OFF_V + s * C::V_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < NL_BN / 16; ++kk)
          if (!NL_DBG(4))
            tc_mma_ts(tmem_o, p_tmem + kk * 8,
                      pos_rows ? smem_desc_sw128_mn(v_addr + kk * (16 * 128), NL_BN * 128, 1024)  // 16 key rows on
                               : smem_desc_sw128(v_addr + kk * 32),
                      idesc_pv, (jt > 0 || kk > 0) ? 1u : 0u);
        tc_commit(&bars[B_PVDONE + jt % NL_EVT]);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax / correction / epilogue =====
    // thread = (query row, half of the key columns); warps w and w + 4 share TMEM lane quadrant w & 3
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int row = quad * 32 + lane;  // 0..127
    const uint32_t pair_bar = 1 + quad;  // named barrier of the two warps of a quadrant
    float m_ref = 0.f, l = 0.f;          // m_ref in units of the raw logits
    {  // this row's half of theta as bf16 pairs into tensor memory: the A operand of every Q K^T.  theta is
      // read in place (Qp = theta, `OutT` elements in the layout `out_cl` says): thread = (query row, half
      // of the channels); rows past HW are zero
      const int qrow = qb * NL_BM + row;
      const bool qok = qrow < HW;
      constexpr int QC = D / 4;  // 32-bit columns (bf16 pairs) per thread
      const OutT* th = reinterpret_cast<const OutT*>(Qp);
#pragma unroll
      for (int c0 = 0; c0 < QC; c0 += 16) {
        uint32_t qv[16];
        if (!qok) {
#pragma unroll
          for (int v = 0; v < 16; ++v) qv[v] = 0u;
        } else if (out_cl) {  // [B][HW][D]: this thread's 32 channels are contiguous
          const OutT* src = th + ((size_t)b * HW + qrow) * D + half * (D / 2) + c0 * 2;
          if constexpr (sizeof(OutT) == 2) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + v);
              qv[4 * v] = t.x; qv[4 * v + 1] = t.y; qv[4 * v + 2] = t.z; qv[4 * v + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(src) + v);
              qv[2 * v] = pack_bf16(t.x, t.y);
              qv[2 * v + 1] = pack_bf16(t.z, t.w);
            }
          }
        } else {  // [B][D][HW]: lanes = consecutive positions, channels HW apart
          const OutT* src = th + ((size_t)b * D + half * (D / 2) + c0 * 2) * HW + qrow;
#pragma unroll
          for (int v = 0; v < 16; ++v)
            qv[v] = pack_bf16(nl_to_float(__ldg(src + (size_t)(2 * v) * HW)), nl_to_float(__ldg(src + (size_t)(2 * v + 1) * HW)));
        }
        tmem_st16(tmem_q + lane_base + half * QC + c0, qv);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_QFULL]);
    }
#ifdef ARFE_PROFILE
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (int it = 0; it < n_it; ++it) {
      const int buf = it & 1, sbuf = it % C::SBUF;
      nl_wait(&bars[B_QKDONE + it % NL_EVT], (uint32_t)(it / NL_EVT) & 1u);
      tc_fence_after();
      NL_T(0);
      if (NL_DBG(2)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_PFULL + sbuf]);
        continue;
      }
      uint32_t sr[32];
      if (!NL_DBG(128)) {
        tmem_ld32(tmem_s + lane_base + sbuf * NL_BN + half * 32, sr);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) sr[c] = __float_as_uint((float)(c + it + lane) * 0.01f);
      }
      NL_T(1);
      const int nvalid = HW - (kb_lo + it) * NL_BN - half * 32;  // columns of this thread that are real keys
      if (nvalid < 32) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c >= nvalid) sr[c] = __float_as_uint(-INFINITY);
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 32; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(sr[c]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // row maximum over both halves (double-buffered exchange: one pair barrier per step)
      if (!NL_DBG(16)) {
        xch[(buf * 2 + half) * NL_BM + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, xch[(buf * 2 + (half ^ 1)) * NL_BM + row]);
      }
      NL_T(2);
      if (it == 0) {
        m_ref = mx;
      } else {
        const bool grow = (mx - m_ref) * sl2 > NL_RESCALE;
        if (__any_sync(0xffffffffu, grow)) {
          // O must be quiescent: P V of the previous step done (it was issued after this step's Q K^T)
          nl_wait(&bars[B_PVDONE + (it - 1) % NL_EVT], (uint32_t)((it - 1) / NL_EVT) & 1u);
          tc_fence_after();
          const float m_new = grow ? mx : m_ref;
          const float alpha = fast_exp2((m_ref - m_new) * sl2);
          l *= alpha;
          m_ref = m_new;
#pragma unroll 1
          for (int ch = 0; ch < D / 64; ++ch) {  // this warp's half of the channels
            uint32_t o[32];
            const uint32_t addr = tmem_o + lane_base + half * (D / 2) + ch * 32;
            tmem_ld32(addr, o);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st32(addr, o);
          }
          tmem_wait_st();
        }
      }
      // P = exp2((s - m_ref) * sl2) as bf16 pairs over this row's S columns in tensor memory (both
      // halves have their S values in registers since the pair barrier): the A operand of P V
      const float neg_m = -m_ref * sl2;
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pw[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float p0 = fmaf(__uint_as_float(sr[2 * e]), sl2, neg_m), p1 = fmaf(__uint_as_float(sr[2 * e + 1]), sl2, neg_m);
        if (!NL_DBG(32)) {
          p0 = fast_exp2(p0);
          p1 = fast_exp2(p1);
        }
        ls4[e & 3] += p0 + p1;
        pw[e] = pack_bf16(p0, p1);
      }
      l += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
      NL_T(3);
      if (!NL_DBG(64)) {
        tmem_st16(tmem_s + lane_base + sbuf * NL_BN + half * 16, pw);
        tmem_wait_st();
      } else if (pw[3] == 0x12345678u) {
        l += 1.f;
      }
      NL_T(4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL + sbuf]);
      NL_T(5);
    }
#ifdef ARFE_PROFILE
    if (NL_DBG(65536) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0)
      printf("softmax warp 0, cycles per step over %d steps: wait S %lld | ld S %lld | max + exchange %lld | rescale check + exp + pack %lld | st P %lld | fence + arrive %lld\n",
             n_it, tacc[0] / n_it, tacc[1] / n_it, tacc[2] / n_it, tacc[3] / n_it, tacc[4] / n_it, tacc[5] / n_it);
#endif
    // epilogue: total row sum, then this warp's half of the channels
    xch[((n_it & 1) * 2 + half) * NL_BM + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    l += xch[((n_it & 1) * 2 + (half ^ 1)) * NL_BM + row];
    nl_wait(&bars[B_PVDONE + (n_it - 1) % NL_EVT], (uint32_t)((n_it - 1) / NL_EVT) & 1u);
    tc_fence_after();
    const int p = qb * NL_BM + row;
    const int NB = B * nqb, blk = b * nqb + qb;
    // all MMAs and with them all tile copies are complete: the K ring is free; per warp a 32 x 33 float
    // staging tile turns "lane = row" into "lane = channel" for channels-last output
    float* stage = reinterpret_cast<float*>(smem + C::OFF_K) + warp * (32 * 33);
    // 32 channels d0.. of this thread's row, scaled: lanes are consecutive positions
    auto write_final = [&](const float (&v)[32], int d0) {
      if (!out_cl) {  // [B][D][HW]: 128 contiguous bytes per channel and warp
        if (p < HW) {
          OutT* dst = y + ((size_t)b * D + d0) * HW + p;
#pragma unroll
          for (int c = 0; c < 32; ++c) nl_store1(dst + (size_t)c * HW, v[c]);
        }
      } else {  // [B][HW][D]: transpose the 32 x 32 tile through shared memory
#pragma unroll
        for (int c = 0; c < 32; ++c) stage[lane * 33 + c] = v[c];
        __syncwarp();
        const int p0 = qb * NL_BM + quad * 32;
#pragma unroll 4
        for (int r = 0; r < 32; ++r)
          if (p0 + r < HW) nl_store1(y + ((size_t)b * HW + p0 + r) * D + d0 + lane, stage[r * 33 + lane]);
        __syncwarp();
      }
    };
    if (nsplit == 1) {
      const float inv = 1.f / l;
#pragma unroll 1
      for (int ch = 0; ch < D / 64; ++ch) {
        uint32_t o[32];
        float v[32];
        const int d0 = half * (D / 2) + ch * 32;
        tmem_ld32(tmem_o + lane_base + d0, o);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(o[c]) * inv;
        write_final(v, d0);
      }
    } else {
      // Key range split: every CTA leaves its (O, max, sum) in the workspace -- O column-major inside the
      // query block, so that lanes = rows store contiguously -- and the CTA of a query block that
      // arrives LAST (a counter, nobody waits) merges:  y = sum_z O_z 2^(m_z - M) / sum_z l_z 2^(m_z - M)
      const float m_own = m_ref * sl2;  // log2 units
      if (half == 0) {
        part_ml[(((size_t)z * NB + blk) * NL_BM + row) * 2 + 0] = m_own;
        part_ml[(((size_t)z * NB + blk) * NL_BM + row) * 2 + 1] = l;
      }
#pragma unroll 1
      for (int ch = 0; ch < D / 64; ++ch) {
        uint32_t o[32];
        const int d0 = half * (D / 2) + ch * 32;
        tmem_ld32(tmem_o + lane_base + d0, o);
        tmem_wait_ld();
        float* dst = part_o + (((size_t)z * NB + blk) * D + d0) * NL_BM + row;
#pragma unroll
        for (int c = 0; c < 32; ++c) dst[(size_t)c * NL_BM] = __uint_as_float(o[c]);
      }
      __threadfence();
      int* flag = reinterpret_cast<int*>(xch);
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (tid == 0) flag[0] = atomicAdd(&counters[blk], 1);
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (flag[0] == nsplit - 1) {
        __threadfence();
        float M = m_own;
        for (int zz = 0; zz < nsplit; ++zz)
          if (zz != z) M = fmaxf(M, __ldcg(&part_ml[(((size_t)zz * NB + blk) * NL_BM + row) * 2]));
        const float w_own = exp2f(m_own - M);
        float den = l * w_own;
        for (int zz = 0; zz < nsplit; ++zz)
          if (zz != z) {
            const float* ml = &part_ml[(((size_t)zz * NB + blk) * NL_BM + row) * 2];
            den += __ldcg(ml + 1) * exp2f(__ldcg(ml) - M);
          }
        const float inv = 1.f / den;
#pragma unroll 1
        for (int ch = 0; ch < D / 64; ++ch) {
          uint32_t o[32];
          float v[32];
          const int d0 = half * (D / 2) + ch * 32;
          tmem_ld32(tmem_o + lane_base + d0, o);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(o[c]) * w_own;
          for (int zz = 0; zz < nsplit; ++zz)
            if (zz != z) {
              const float wz = exp2f(__ldcg(&part_ml[(((size_t)zz * NB + blk) * NL_BM + row) * 2]) - M);
              const float* src = part_o + (((size_t)zz * NB + blk) * D + d0) * NL_BM + row;
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = fmaf(__ldcg(src + (size_t)c * NL_BM), wz, v[c]);
            }
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] *= inv;
          write_final(v, d0);
        }
        if (tid == 0) counters[blk] = 0;  // ready for the next call on this workspace
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == NL_SOFTMAX_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Operand packing of phi and g (theta is read in place): fp32 / bf16 -> bf16 tiles of 64 positions in
// the swizzled shared-memory image, straight from global memory (two 16-byte chunks per thread, no
// staging, no transposition -- the MMA descriptors take either orientation):
//   channels-last [B][HW][D]: tile = [D/64 slabs][64 position rows][128 B = 64 channels]
//                             (phi: K-major operand of Q K^T; g: MN-major operand of P V)
//   NCHW          [B][D][HW]: tile = [D channel rows][128 B = 64 positions]
//                             (phi: MN-major operand of Q K^T; g: K-major operand of P V)
template <typename T>
__global__ void __launch_bounds__(256)
nl_pack_kernel(const T* __restrict__ phi, const T* __restrict__ g, uint8_t* __restrict__ Kp, uint8_t* __restrict__ Vp,
               int HW, int D, int in_cl, int nkb, int* __restrict__ counters, int ncounters) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)  // arrival counters of the key-range merge
    for (int i = threadIdx.x; i < ncounters; i += blockDim.x) counters[i] = 0;
  const int which = blockIdx.z & 1, b = blockIdx.z >> 1;  // 0: phi, 1: g
  const int pb = blockIdx.x, slab = blockIdx.y;
  const T* src = which == 0 ? phi : g;
  uint8_t* tile = (which == 0 ? Kp : Vp) + ((size_t)b * nkb + pb) * ((size_t)NL_BN * D * 2);
  const int p0 = pb * 64, c0 = slab * 64;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q = threadIdx.x + 256 * i, r = q >> 3, c = q & 7;
    float v[8];
    uint8_t* dst;
    if (in_cl) {  // row r = position, chunk c = 8 channels
      const int prow = p0 + r;
      if (prow < HW) {
        const T* sp = src + ((size_t)b * HW + prow) * D + c0 + c * 8;
        if constexpr (sizeof(T) == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(sp)), bb = __ldg(reinterpret_cast<const float4*>(sp) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bb.x; v[5] = bb.y; v[6] = bb.z; v[7] = bb.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = nl_to_float(sp[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
      dst = tile + (size_t)slab * (NL_BN * 128) + r * 128 + ((c ^ (r & 7)) * 16);
    } else {  // row d = channel, chunk c = 8 positions
      const int d = c0 + r;
      const T* sp = src + ((size_t)b * D + d) * HW;
      const int pp0 = p0 + c * 8;
      if (sizeof(T) == 4 && (HW & 3) == 0 && pp0 + 8 <= HW) {  // rows 16-byte aligned: two vector loads
        const float4 a = __ldg(reinterpret_cast<const float4*>(sp + pp0)), bb = __ldg(reinterpret_cast<const float4*>(sp + pp0) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bb.x; v[5] = bb.y; v[6] = bb.z; v[7] = bb.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = pp0 + e < HW ? nl_to_float(__ldg(sp + pp0 + e)) : 0.f;
      }
      dst = tile + (size_t)d * 128 + ((c ^ (d & 7)) * 16);
    }
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// ------------------------------------------------------------------------------------------------
// Row pass of the attention BACKWARD (the GEMMs around it stay library calls): from the logits S and
// dP = dY . g^T of one query row,
//   P = softmax(scale * S),  delta = sum_q P dP,  dS = scale * P * (dP - delta)
// P and dS leave as bf16 (the operands of the three GEMMs that follow); S and dP are read once.
__global__ void __launch_bounds__(256)
nl_bwd_rows_kernel(const float* __restrict__ S, const float* __restrict__ dP, __nv_bfloat16* __restrict__ Pb,
                   __nv_bfloat16* __restrict__ dSb, int n, float scale) {
  extern __shared__ float prow[];  // the row's exponentials
  __shared__ float red[8];
  const size_t base = (size_t)blockIdx.x * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float sl2 = scale * 1.4426950408889634f;
  auto block_reduce = [&](float v, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
  };
  float mx = -INFINITY;
  for (int i = tid; i < n; i += 256) {
    const float v = __ldg(S + base + i);
    prow[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_reduce(mx, true);
  float sum = 0.f, dot = 0.f;
  for (int i = tid; i < n; i += 256) {
    const float e = exp2f((prow[i] - mx) * sl2);
    prow[i] = e;
    sum += e;
    dot = fmaf(e, __ldg(dP + base + i), dot);
  }
  sum = block_reduce(sum, false);
  dot = block_reduce(dot, false);
  const float inv = 1.f / sum, delta = dot * inv;
  for (int i = tid; i < n; i += 256) {
    const float pv = prow[i] * inv;
    Pb[base + i] = __float2bfloat16(pv);
    dSb[base + i] = __float2bfloat16(scale * pv * (__ldg(dP + base + i) - delta));
  }
}

struct NlLayout {
  int nqb, nkb, rows_pad;
  size_t q_bytes, k_bytes, v_bytes, part_o_bytes, part_ml_bytes;
  size_t off_k, off_v, off_po, off_pml, off_cnt, total;
};
NlLayout nl_layout(int B, int HW, int D, int nsplit) {
  NlLayout a;
  a.nqb = (HW + NL_BM - 1) / NL_BM;
  a.nkb = (HW + NL_BN - 1) / NL_BN;
  a.rows_pad = a.nqb * NL_BM;
  a.q_bytes = 0;  // theta is read in place
  a.k_bytes = (size_t)B * a.nkb * NL_BN * D * 2;
  a.v_bytes = a.k_bytes;
  a.part_o_bytes = nsplit > 1 ? (size_t)nsplit * B * a.rows_pad * D * 4 : 0;
  a.part_ml_bytes = nsplit > 1 ? (size_t)nsplit * B * a.rows_pad * 2 * 4 : 0;
  a.off_k = a.q_bytes;
  a.off_v = a.off_k + a.k_bytes;
  a.off_po = a.off_v + a.v_bytes;
  a.off_pml = a.off_po + a.part_o_bytes;
  a.off_cnt = a.off_pml + a.part_ml_bytes;  // one arrival counter per query block (zeroed by the pack kernel)
  a.total = a.off_cnt + (((size_t)B * a.nqb * 4 + 255) & ~(size_t)255);
  return a;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// [B][HW][D] bf16 rows -> boxes of 64 channels x 64 positions, 128-byte swizzle, zero fill out of bounds
bool make_tile_map(CUtensorMap* tm, const void* base, int B, int HW, int D) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)HW, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)HW * D * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)NL_BN, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D, typename OutT, bool TM, bool POS>
cudaError_t nl_launch_attn(const NlLayout& a, const uint8_t* q, const uint8_t* k, const uint8_t* v, const CUtensorMap& tk,
                           const CUtensorMap& tv, uint8_t* ws, OutT* y, int B, int HW, int nsplit, float sl2,
                           int out_cl, cudaStream_t stream) {
  using C = NlCfg<D>;
  cudaError_t e = cudaFuncSetAttribute(nl_attn_kernel<D, OutT, TM, POS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return e;
  nl_attn_kernel<D, OutT, TM, POS><<<dim3(a.nqb, B, nsplit), NL_THREADS, C::SMEM, stream>>>(
      q, k, v, tk, tv, y, reinterpret_cast<float*>(ws + a.off_po), reinterpret_cast<float*>(ws + a.off_pml),
      reinterpret_cast<int*>(ws + a.off_cnt), HW, a.nqb, a.nkb, nsplit, sl2, out_cl, ARFE_KNOB_ENV("ARFE_NL_DBG", 0));
  return cudaGetLastError();
}

template <typename T>
cudaError_t nl_run(const void* theta, const void* phi, const void* g, void* y, int B, int HW, int D, int in_cl,
                   float scale, void* workspace, int nsplit, cudaStream_t stream) {
  const NlLayout a = nl_layout(B, HW, D, nsplit);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const float sl2 = scale * 1.4426950408889634f;
  T* yo = static_cast<T*>(y);
  cudaError_t e;
  CUtensorMap tk, tv;
  memset(&tk, 0, sizeof(tk));
  memset(&tv, 0, sizeof(tv));
  // bf16 channels-last inputs are already what the kernel reads: theta rows directly, phi / g tiles through
  // tensor maps; only the arrival counters of the key-range merge need a reset
  if (sizeof(T) == 2 && in_cl && (reinterpret_cast<uintptr_t>(phi) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(theta) & 15) == 0 && make_tile_map(&tk, phi, B, HW, D) && make_tile_map(&tv, g, B, HW, D)) {
    if (nsplit > 1 && (e = cudaMemsetAsync(ws + a.off_cnt, 0, (size_t)B * a.nqb * 4, stream)) != cudaSuccess) return e;
    const uint8_t* q = static_cast<const uint8_t*>(theta);
    switch (D) {
      case 64: return nl_launch_attn<64, T, true, true>(a, q, nullptr, nullptr, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream);
      case 128: return nl_launch_attn<128, T, true, true>(a, q, nullptr, nullptr, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream);
      case 256: return nl_launch_attn<256, T, true, true>(a, q, nullptr, nullptr, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream);
      default: return cudaErrorNotSupported;
    }
  }
  nl_pack_kernel<T><<<dim3(a.nkb, D / 64, 2 * B), 256, 0, stream>>>(
      static_cast<const T*>(phi), static_cast<const T*>(g), ws + a.off_k, ws + a.off_v, HW, D, in_cl, a.nkb,
      reinterpret_cast<int*>(ws + a.off_cnt), B * a.nqb);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const uint8_t* qsrc = static_cast<const uint8_t*>(theta);
  switch (D) {
    case 64: return in_cl ? nl_launch_attn<64, T, false, true>(a, qsrc, ws + a.off_k, ws + a.off_v, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream)
                         : nl_launch_attn<64, T, false, false>(a, qsrc, ws + a.off_k, ws + a.off_v, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream);
    case 128: return in_cl ? nl_launch_attn<128, T, false, true>(a, qsrc, ws + a.off_k, ws + a.off_v, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream)
                         : nl_launch_attn<128, T, false, false>(a, qsrc, ws + a.off_k, ws + a.off_v, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream);
    case 256: return in_cl ? nl_launch_attn<256, T, false, true>(a, qsrc, ws + a.off_k, ws + a.off_v, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream)
                         : nl_launch_attn<256, T, false, false>(a, qsrc, ws + a.off_k, ws + a.off_v, tk, tv, ws, yo, B, HW, nsplit, sl2, in_cl, stream);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace

cudaError_t launch_nonlocal_backward_rows(const float* S, const float* dP, void* Pb, void* dSb, long long rows, int n,
                                          float scale, cudaStream_t stream) {
  if (rows <= 0 || n <= 0) return cudaSuccess;
  const size_t smem = (size_t)n * sizeof(float);
  if (smem > 200 * 1024 || rows > 0x7fffffffll) return cudaErrorNotSupported;
  cudaError_t e;
  if (smem > 48 * 1024 &&
      (e = cudaFuncSetAttribute(nl_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return e;
  nl_bwd_rows_kernel<<<(unsigned)rows, 256, smem, stream>>>(S, dP, static_cast<__nv_bfloat16*>(Pb),
                                                           static_cast<__nv_bfloat16*>(dSb), n, scale);
  return cudaGetLastError();
}

int nonlocal_default_split(int B, int HW) {
  if (B < 1 || HW < 1) return 1;
  const int nqb = (HW + NL_BM - 1) / NL_BM, nkb = (HW + NL_BN - 1) / NL_BN;
  int s = sm_count() / (nqb * B);
  if (s < 1) s = 1;
  if (s > 4) s = 4;
  if (s > nkb) s = nkb;
  return s;
}

size_t nonlocal_workspace_bytes(int B, int HW, int D, int nsplit) { return nl_layout(B, HW, D, nsplit).total; }

cudaError_t launch_nonlocal_attention(const void* theta, const void* phi, const void* g, void* y, int B, int HW, int D,
                                      int dtype, int in_cl, float scale, void* workspace, int nsplit,
                                      cudaStream_t stream) {
  if (D != 64 && D != 128 && D != 256) return cudaErrorNotSupported;
  return dtype == 0 ? nl_run<float>(theta, phi, g, y, B, HW, D, in_cl, scale, workspace, nsplit, stream)
                    : nl_run<__nv_bfloat16>(theta, phi, g, y, B, HW, D, in_cl, scale, workspace, nsplit, stream);
}

}  // namespace arfe
