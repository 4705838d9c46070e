// tcgen05 engine: shared-basis sample paths as a 3xTF32 tensor-core contraction whose
// A operand (the RFF / kernel features) is generated on the fly and never leaves the SM.
//
// One persistent CTA per SM, 24 warps, warp-specialised:
//   warp 21     B loader   : bulk-async copies (cp.async.bulk, TMA unit) of 16 KB blocks of the
//                            hi/lo-split [W;V] operand, pre-laid-out in UMMA canonical K-major
//                            form by tensor_pack(), into a 4-stage shared-memory ring
//   warp 20     MMA issuer : one thread issues tcgen05.mma kind::tf32, A from TENSOR MEMORY,
//                            B from shared memory, fp32 accumulators in tensor memory
//   warp 22     TMEM allocator
//   warps 0-15  producers  : lane = candidate row (x in registers).  Per 32-feature block each
//                            thread evaluates 8 features (FFMA projection + MUFU cos, or the
//                            Matern/RBF profile), splits them into TF32 hi/lo and writes them
//                            straight into the A stage in tensor memory (tcgen05.st)
//   warps 16-19 epilogue   : tcgen05.ld the finished 128 x 64 tile, stage it in shared memory
//                            and run the same fused reduction as the SIMT kernels (level-set
//                            ballots / top-k filter / values dump) while the next tile's MMAs run
//
// Split precision, fp32 accumulate: f = Phi_1*W_1 + Phi_1*W_2 + Phi_2*W_1, three M=128, N=64 MMAs
// per k-step into the same 64 accumulator columns.  Two operand formats:
//   3xTF32  (PSBAX_ENGINE_TENSOR)      x = hi + lo, hi = TF32(x): ~2^-21 per product; K=8 per MMA
//   BF16x3  (PSBAX_ENGINE_TENSOR_BF16) x = x1 + x2, x1 = bf16(x), x2 = bf16(x - x1): <= 3*2^-18 per
//           product (6x inside the 1e-4 tolerance); K=16 per MMA at the same cycles, half the
//           operand bytes in shared and tensor memory -> 4 A stages instead of 2.
//
// A CTA pass covers TWO 128-row tiles (256 candidates) against every [W;V] block: each table
// row read from shared memory and each B block serve two evaluations, and the per-block
// synchronisation is amortised over 24 MMAs.
//
// Tensor-memory columns (512 allocated):
//   D[buf][tile]     : (buf*2 + tile)*64               buf, tile in {0,1}  (accumulators, double-buffered)
//   A[stage][tile]   : 256 + (stage*2 + tile)*ACT      first split at +0, second at +ACT/2
//                      ACT = 64 (TF32, 2 stages) or 32 (BF16 pairs, 4 stages)
#pragma once
#include "sp_common.cuh"
#include <cstdio>
#include <type_traits>

#include "sp_simt.cuh"

namespace psbax {

constexpr int TC_MAX_D = 128;           // d <= 16: x in registers; larger d: x tile in shared memory
constexpr int TC_MAX_D_REG = 16;
constexpr int TC_MAX_S = 1 << 20;  // any S: processed in chunks of 64 samples
#ifndef PSBAX_TC_NPROD
#define PSBAX_TC_NPROD 16
#endif
#ifndef PSBAX_TC_FT3
#define PSBAX_TC_FT3 1  // d <= 2: a second, dense copy of the feature table (12-byte rows): six LDS.128 per eight features
#endif
#ifndef PSBAX_TC_SLEEP_NS
#define PSBAX_TC_SLEEP_NS 2000  // real sleeps (ns) between the fallback probes of the roles with slack: loader (b_empty), epilogue (d_full)
#endif
#ifndef PSBAX_EXP_NOLDS
#define PSBAX_EXP_NOLDS 0
#endif
#ifndef PSBAX_EXP_NOCOS
#define PSBAX_EXP_NOCOS 0
#endif
#ifndef PSBAX_EXP_NOSTTM
#define PSBAX_EXP_NOSTTM 0
#endif
#ifndef PSBAX_EXP_NOEPI
#define PSBAX_EXP_NOEPI 0
#endif
#ifndef PSBAX_EXP_NOCOPY
#define PSBAX_EXP_NOCOPY 0
#endif
#ifndef PSBAX_EXP_FREERUN
#define PSBAX_EXP_FREERUN 0  // what-if build (wrong results): producers alone, no barriers, every other role idle
#endif
#ifndef PSBAX_EXP_NOAEMPTY
#define PSBAX_EXP_NOAEMPTY 0  // what-if build (wrong results): producers never wait for their A stage
#endif
#ifndef PSBAX_EXP_NOMMA
#define PSBAX_EXP_NOMMA 0
#endif
#ifndef PSBAX_SPLIT_FHFMA
#define PSBAX_SPLIT_FHFMA 1  // BF16 residuals by mixed-precision FMA (0: unpack + packed subtract, the round-1 form)
#endif
// Operand-format / block-size configuration.  GEN = generic-d kernel (x tile in shared memory).
template <bool BF, bool GEN>
struct TcCfg {
  // features per block (= per A / B stage).  BF16x3 with x in registers uses 64-feature blocks:
  // same stage footprints as 3xTF32, half the synchronisation per feature (measured 10.0 -> 8.9 ms)
  static constexpr int TKB = (BF && !GEN) ? 64 : 32;
  static constexpr int ESZ = BF ? 2 : 4;                 // operand element bytes
  static constexpr int KMMA = BF ? 16 : 8;               // K per MMA
  static constexpr int BBYTES = 128 * TKB * ESZ;         // B image: 128 rows (split1|split2) x TKB
  static constexpr int ACT = TKB * ESZ * 2 / 4;          // A columns per tile per stage (two splits)
  static constexpr int NSA = 128 / ACT;                  // A ring depth: 256 TMEM columns / (2 tiles * ACT)
  static constexpr int NSB = GEN ? 4 : (BBYTES == 16384 ? 4 : 8);  // B ring depth
  static constexpr int KSTEPS = TKB / KMMA;              // MMAs along k per block
  static constexpr int KCOLS = 8;                        // A columns per k-step
  static constexpr int B_KSTEP = 256;                    // bytes: two 128-byte core matrices per k-step
  static constexpr int B_SBO = TKB * ESZ * 8;            // bytes between 8-row groups
  static constexpr int B_SPLIT2 = 8 * B_SBO;             // rows 64..127
  static constexpr int WCOLS = BF ? 4 : 8;               // A columns per 8 features per split
  // A-stage column layout of a tile.  Default: [split 1: all features][split 2: all features].
  // 64-feature BF16x3 blocks: per k-step (= the 16 features of one producer warp) the two splits
  // sit side by side, [k-step][split][8 columns], so a warp stores its block with ONE 16-column
  // tcgen05.st per tile.
  static constexpr bool INTERLEAVED = BF && !GEN;
  static constexpr int A_KSTEP = INTERLEAVED ? 2 * KCOLS : KCOLS;   // columns between k-steps
  static constexpr int A_SPLIT2 = INTERLEAVED ? KCOLS : ACT / 2;    // offset of split 2
};
__host__ __device__ inline int tc_tkb(bool bf, bool generic) { return (bf && !generic) ? 64 : 32; }
__host__ __device__ inline uint32_t tc_bbytes(bool bf, bool generic) { return 128u * tc_tkb(bf, generic) * (bf ? 2 : 4); }

#ifndef PSBAX_TC_NPROD
#define PSBAX_TC_NPROD 16
#endif
// warp roles: producers first (lane quarter = warp % 4), then the epilogue, the MMA issuer, the
// loader and the TMEM allocator.  (Measured: the order of epilogue and producers does not matter;
// the epilogue must simply be short enough to hide behind a pass — see profiles/r01_tuning_log.md.)
constexpr int TC_PROD_WARP0 = 0, TC_NPROD = PSBAX_TC_NPROD;          // 8 or 16 producer warps
constexpr int TC_EPI_WARP0 = TC_NPROD, TC_NEPI = 4;
constexpr int TC_WARP_MMA = TC_EPI_WARP0 + TC_NEPI, TC_WARP_LOAD = TC_WARP_MMA + 1,
              TC_WARP_ALLOC = TC_WARP_MMA + 2;
constexpr int TC_THREADS = (TC_WARP_MMA + 4) * 32;
static_assert(TC_NPROD == 8 || TC_NPROD == 16, "8 or 16 producer warps");
constexpr int TC_COL_D = 0, TC_COL_A = 256;
constexpr int TC_TILES = 2;                  // row tiles per CTA pass
constexpr uint32_t TC_TMEM_COLS = 512;

inline bool tensor_supported(const psbax_paths_t* p, bool bf16);
inline bool is_tensor_mode(int mode) { return mode == MODE_TENSOR || mode == MODE_TENSOR_BF16; }
// [chunks][nkb] images of BBYTES
inline size_t tensor_pack_floats(const Layout& l) {
  const bool bf = l.mode == MODE_TENSOR_BF16;
  return (size_t)(l.Sp / TN) * (l.nkb_r + l.nkb_k) * (tc_bbytes(bf, l.tc_generic != 0) / 4);
}

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait: the hardware parks the thread for a bounded time and wakes it when the phase
// completes (PSBAX_TC_WAIT_HINT adds an explicit suspend-time hint, in ns).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if defined(PSBAX_TC_WAIT_HINT)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)PSBAX_TC_WAIT_HINT)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// non-blocking probe (no suspend): issued early so that its latency overlaps with compute
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) asm volatile("trap;");
  }
}
// one lane of a converged warp; the compiler then issues uniform-datapath instructions
// (UTCHMMA, UBLKCP, UTCBAR) directly instead of wrapping each in an elect loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// relaxed wait for roles with slack (producers on a 4-stage ring): a failed probe sleeps a few
// tens of ns instead of re-polling at full rate (polls were ~15 % of all issued instructions)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if (++spins > (1u << 24)) asm volatile("trap;");
  }
}
// one lane polls, the warp follows: 32 lanes hitting the same mbarrier word serialise in the
// shared-memory sync unit (measured: a dominant per-block cost with 16 producer warps)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}
// Warp-level wait for roles with many warps: one converged non-blocking probe by all lanes (the
// common case: already complete, no divergence); only if that fails does lane 0 poll while the
// others park at the reconvergence point.  A lane-0-only poll always leaves the warp diverged,
// and the WARPSYNC slow path of the reconvergence costs a few hundred cycles per wait (ncu source
// view of the projection kernel: half of the issuer warp's samples sat in those paths).
__device__ __forceinline__ void mbar_wait_hybrid(uint64_t* bar, uint32_t parity, int lane) {
  // one vote instead of an unconditional reconvergence barrier after a possibly divergent branch
  if (!__all_sync(0xffffffffu, mbar_test(bar, parity))) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
  }
}
// the same with a sleep between the fallback probes, for roles that wait long (the epilogue waits
// most of a pass for d_full; its polls otherwise take issue slots from the producers)
// try_wait with an explicit suspend-time hint (ns): the hardware parks the thread until the phase
// completes or the hint expires, so a long wait costs a handful of instructions instead of a poll
// every ~45 ns (the default limit; __nanosleep(500) measured no longer than that either)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// warp-level wait for roles with slack: one converged probe; on failure lane 0 sleeps NS ns between probes
template <int NS>
__device__ __forceinline__ void mbar_wait_hybrid_nanosleep(uint64_t* bar, uint32_t parity, int lane) {
  if (!__all_sync(0xffffffffu, mbar_test(bar, parity))) {
    if (lane == 0) {
      uint32_t spins = 0;
      while (!mbar_test(bar, parity)) {
        __nanosleep(NS);
        if (++spins > (1u << 24)) asm volatile("trap;");
      }
    }
    __syncwarp();
  }
}
#ifndef PSBAX_AEMPTY_HINT
#define PSBAX_AEMPTY_HINT 0  // measured: 0 (plain try_wait), 300, 1000, 5000 ns all 7.82 ms
#endif
template <int NS>
__device__ __forceinline__ void mbar_wait_hybrid_sleep(uint64_t* bar, uint32_t parity, int lane) {
  if (!__all_sync(0xffffffffu, mbar_test(bar, parity))) {
    if (lane == 0) {
      uint32_t spins = 0;
      while (!(NS > 0 ? mbar_try_wait_hint(bar, parity, NS) : mbar_try_wait(bar, parity))) {
        if (++spins > (1u << 24)) asm volatile("trap;");
      }
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, cta_group::1
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
// packed fp32x2 arithmetic (FFMA2 on sm_100): one instruction for the two row tiles of a thread
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// {hi16, lo16} = {bf16(a), bf16(b)}, round to nearest even
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b));
  return d;
}
// residuals of a packed bf16 pair against its fp32 sources, one mixed-precision FMA each (FHFMA.BF16 on
// sm_100: the bf16 half is an operand selector, so neither the unpack nor a packed subtract is needed):
// r_lo = v_lo - bf16_lo(p), r_hi = v_hi - bf16_hi(p)
__device__ __forceinline__ void bf16_residual2(uint32_t p, float v_lo, float v_hi, float& r_lo, float& r_hi) {
  asm("{\n\t.reg .b16 l, h, m;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "mov.b16 m, 0xBF80;\n\t"  // -1.0
      "fma.rn.f32.bf16 %0, l, m, %3;\n\t"
      "fma.rn.f32.bf16 %1, h, m, %4;\n\t}"
      : "=f"(r_lo), "=f"(r_hi)
      : "r"(p), "f"(v_lo), "f"(v_hi));
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): 8-row x 16-byte core
// matrices stored contiguously (128 B); LBO = byte distance between the two core matrices
// that are adjacent in K, SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor: (TF32 x TF32 | BF16 x BF16) -> F32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// clock trace (compile with -DPSBAX_TC_TRACE=1): role r (0 = MMA, 1 = first producer warp,
// 2 = first epilogue warp, 3 = loader), CTA 0 only.  Compiled out of the production build.
// PSBAX_TC_TRACE=2: the four producer warps of lane quarter 0 (one scheduler) as roles 0..3 instead.
#if defined(PSBAX_TC_TRACE) && PSBAX_TC_TRACE
#define TC_TRACE_RAW(role, it, slot)                                                     \
  do {                                                                                   \
    if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (it) < 128)          \
      a.trace[((role) * 128 + (it)) * 8 + (slot)] = clock64();                           \
  } while (0)
#if PSBAX_TC_TRACE == 2
#define TC_TRACE(role, it, slot) do { } while (0)
#define TC_TRACE_P(it, slot) TC_TRACE_RAW(trole, it, slot)
#else
#define TC_TRACE(role, it, slot) TC_TRACE_RAW(role, it, slot)
#define TC_TRACE_P(it, slot) TC_TRACE_RAW(1, it, slot)
#endif
#else
#define TC_TRACE(role, it, slot) do { } while (0)
#define TC_TRACE_P(it, slot) do { } while (0)
#endif

struct TcSmem {
  // byte offsets into dynamic shared memory.  `b`: ring of NSB stages of `stage` bytes: the B
  // image and, for generic d, the 32-row block of the feature table behind it.
  uint32_t b, stage, nsb, ft, ft3, x, out, state, il, bars, total;
  bool generic;
};
__host__ __device__ inline TcSmem tc_smem_layout_impl(const Layout& l, int epi, int k, bool generic) {
  const bool bf = l.mode == MODE_TENSOR_BF16;
  TcSmem s{};
  s.generic = generic;
  const uint32_t tkb = tc_tkb(bf, generic), bbytes = tc_bbytes(bf, generic);
  s.nsb = generic ? 4u : (bbytes == 16384u ? 4u : 8u);
  s.stage = bbytes + (generic ? tkb * l.OS * 4 : 0u);
  Layout lb = l;  // rows of the (resident) feature table for this block size
  lb.Lp_ft = l.Lp;
  lb.tkb = (int)tkb;
  layout_blocks(lb);
  const uint32_t kp = (uint32_t)lb.Kp;
  uint32_t o = 0;
  s.b = o;     o += s.nsb * s.stage;
  s.ft = o;    o += generic ? 0u : kp * l.OS * 4;                         // resident table (small d)
  s.ft3 = o;   o += (PSBAX_TC_FT3 && !generic && l.dp4 == 2) ? (kp * 12u + 15u) / 16u * 16u : 0u;  // dense (omega_1, omega_2, b) rows
  s.x = o;     o += generic ? (uint32_t)TC_TILES * l.dp4 * TM * 4 : 0u;   // x tile [i][row] of (tile 0, tile 1) pairs
  s.out = o;   o += TN * OUT_LD * 4;
  s.state = o; o += (uint32_t)round_up((int)epi_smem_floats(epi, k), 4) * 4;
  s.il = o;    o += (uint32_t)round_up(l.dp4, 4) * 4;
  o = (o + 15) / 16 * 16;
  s.bars = o;  o += 40 * 8;
  s.total = o;
  return s;
}
constexpr uint32_t TC_SMEM_MAX = 227 * 1024;
// Shape-only decision (the operand pack depends on it): resident feature table + x in registers
// when d <= 16 and everything fits shared memory even with the largest reduction state, else the
// generic kernel.  Needs l.mode, l.dp4, l.OS, l.Lp, l.np_.
__host__ __device__ inline bool tc_choose_generic(const Layout& l) {
  if (l.dp4 > TC_MAX_D_REG) return true;
  return tc_smem_layout_impl(l, EPI_TOPK, KMAX, false).total > TC_SMEM_MAX;
}
__host__ __device__ inline TcSmem tc_smem_layout(const Layout& l, int epi, int k) {
  return tc_smem_layout_impl(l, epi, k, l.tc_generic != 0);
}

// CTA c of gx handles the contiguous range [t0, t1) of passes
__host__ __device__ inline void tc_tile_range(int n, int gx, int c, int* t0, int* t1) {
  const int base = n / gx, rem = n % gx;
  *t0 = c * base + (c < rem ? c : rem);
  *t1 = *t0 + base + (c < rem ? 1 : 0);
}

// shape check incl. shared memory with the largest reduction state (top-k, k = KMAX)
inline bool tensor_supported(const psbax_paths_t* p, bool bf16) {
  const bool lin = p->kernel == PSBAX_KERNEL_LINEAR;
  if (!(p->G == 1 && p->S >= 1 && (p->L > 0 || p->n > 0) && p->d <= TC_MAX_D)) return false;
  Layout l{};
  l.mode = bf16 ? MODE_TENSOR_BF16 : MODE_TENSOR;
  l.dp4 = (p->d <= 2) ? 2 : round_up(p->d, 4);
  l.OS = round_up(l.dp4 + 1, 4);
  l.Lp = round_up(p->L, 8) + (lin ? round_up(p->n, 8) : 0);
  l.np_ = lin ? 0 : round_up(p->n, 8);
  const bool generic = tc_choose_generic(l);
  return tc_smem_layout_impl(l, EPI_TOPK, KMAX, generic).total <= TC_SMEM_MAX;
}

inline int tensor_parts(const Layout& l, long long N, int sms) {
  const long long nt = (N + TC_TILES * TM - 1) / (TC_TILES * TM);  // passes of 256 rows
  const int gy = l.Sp / TN;
  // one CTA per SM and ONE wave: gx * gy <= #SM (rounding up put 160 CTAs on 148 SMs for 16 sample
  // chunks: a second wave of 12 CTAs doubled the time)
  long long gx = sms / gy;
  if (gx > nt) gx = nt;
  if (gx < 1) gx = 1;
  return (int)gx;
}

// ---------------------------------------------------------------------------------
// operand pack: the two splits of [W;V] in UMMA canonical (no-swizzle, K-major) layout.
//   image(cy, kb)[row r in 0..127][k in 0..31]:  r < 64: split1(wv[kb*32+k][cy*64+r]),
//                                                r >= 64: split2(wv[kb*32+k][cy*64+r-64])
//   core matrix = 8 rows x 16 bytes, stored contiguously (128 B); K-adjacent core matrices 128 B
//   apart (LBO); 8-row groups SBO apart.  Element offset inside an image:
//     TF32: (r/8)*256 + (k/4)*32 + (r%8)*4 + (k%4)      (floats)
//     BF16: (r/8)*256 + (k/8)*64 + (r%8)*8 + (k%8)      (bf16)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t bf16_rn_bits(float x) { return (uint16_t)(cvt_bf16x2(x, x) & 0xffffu); }

#ifdef PSBAX_TU_API  // the pack kernel lives in the API translation unit only
__global__ void k_tensor_pack(Layout lay, float* __restrict__ pack) {
  const float* __restrict__ wv = pack + lay.off_wv;
  const bool bf = lay.mode == MODE_TENSOR_BF16;
  const int tkb = lay.tkb;
  const int nkb = lay.nkb_r + lay.nkb_k;
  const int wpi = bf ? 64 * tkb : 128 * tkb;  // 4-byte words per image (BBYTES / 4)
  const size_t total = (size_t)(lay.Sp / TN) * nkb * wpi;
  uint32_t* __restrict__ tc = reinterpret_cast<uint32_t*>(pack + lay.off_tc);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int within = (int)(idx % wpi);
    const size_t img = idx / wpi;
    const int kb = (int)(img % nkb), cy = (int)(img / nkb);
    if (bf && kb < lay.nkb_r) {
      // bf16 image of tkb features: word = elements (2 within, 2 within + 1), adjacent in k
      const int group_elems = 8 * tkb;
      uint32_t word = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int el = 2 * within + e;
        const int rg = el / group_elems, rem = el % group_elems;
        const int kc = rem / 64, r8 = (rem >> 3) & 7, ke = rem & 7;
        const int r = rg * 8 + r8, k = kc * 8 + ke;
        const float x = wv[(size_t)(kb * tkb + k) * lay.Sp + cy * TN + (r & 63)];
        const uint16_t b1 = bf16_rn_bits(x);
        const float x1 = __uint_as_float((uint32_t)b1 << 16);
        const uint16_t bits = (r < 64) ? b1 : bf16_rn_bits(x - x1);
        word |= (uint32_t)bits << (16 * e);
      }
      tc[idx] = word;
    } else {
      // tf32 image: kf features (the whole block in the 3xTF32 engine, an exact-kernel block of
      // tkb / 2 rows in the BF16x3 engine)
      const int kf = bf ? tkb / 2 : tkb;
      const int k0 = bf ? lay.kk0 + (kb - lay.nkb_r) * kf : kb * tkb;
      const int group_elems = 8 * kf;
      const int rg = within / group_elems, rem = within % group_elems;
      const int kc = rem / 32, r8 = (rem >> 2) & 7, ke = rem & 3;
      const int r = rg * 8 + r8, k = kc * 4 + ke;
      const float x = wv[(size_t)(k0 + k) * lay.Sp + cy * TN + (r & 63)];
      const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);  // round to TF32
      tc[idx] = __float_as_uint((r < 64) ? hi : (x - hi));
    }
  }
}

inline int tensor_pack(const Layout& l, float* pack, cudaStream_t st) {
  k_tensor_pack<<<296, 256, 0, st>>>(l, pack);
  return cudaGetLastError() == cudaSuccess ? PSBAX_OK : PSBAX_E_CUDA;
}

#endif  // PSBAX_TU_API

// ---------------------------------------------------------------------------------
// Level-set epilogue with register-resident state.  The epilogue warps are on the critical
// path of the tensor engine (4 warps serve a 256-row pass), so: lane = row, every lane keeps its
// own running max per sample, the cross-lane arg-max reduction happens once per kernel, the
// 16 x 4 (sample, row-group) steps are fully unrolled (independent chains), and the four mask
// words of a sample leave as one 16-byte store.
// ---------------------------------------------------------------------------------
template <int NW>
struct LsRegState {
  static constexpr int SPW = TN / NW;  // samples per warp
  float bv[SPW];
  int bi[SPW];
  unsigned int cnt[SPW];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int q = 0; q < SPW; ++q) { bv[q] = -CUDART_INF_F; bi[q] = 0x7fffffff; cnt[q] = 0u; }
  }
};

template <int NW, bool FULL>
__device__ __forceinline__ void epilogue_levelset_reg_impl(const KArgs& a, const float* sOut, LsRegState<NW>& st,
                                                           int row0, int s0, int tid) {
  constexpr int SPW = TN / NW;
  const int warp = tid >> 5, lane = tid & 31;
  const long long N = a.N;
  const bool wr = a.mask != nullptr && lane < TM / 32 && (FULL || (long long)row0 + lane * 32 < N);
  // Everything about the mask store that does not depend on the sample is formed once per tile: lane g < 4 owns
  // the word of row group g, the samples of this warp are mask_ld words apart (the pointer walks).  Under the
  // board's power cap the kernel is energy-bound and this epilogue was 13 % of its executed instructions.
  unsigned int* mp = a.mask + ((long long)(s0 + warp * SPW) * a.mask_ld + (row0 >> 5) + lane);
  const int nvalid = a.lay.S - (s0 + warp * SPW);  // samples of this warp that exist (padding samples evaluate to c0)
  unsigned int pool = 0u;  // lane g < 4: OR over this warp's samples of row group g's word
#pragma unroll
  for (int q = 0; q < SPW; ++q) {
    const int sl = warp * SPW + q;
    unsigned int mine = 0u;
#pragma unroll
    for (int g = 0; g < TM / 32; ++g) {
      const int r = g * 32 + lane;
      const int row = row0 + r;
      const bool valid = FULL || row < N;  // FULL: every row of the tile exists (all tiles but the last)
      const float y = fmaf(sOut[sl * OUT_LD + r], a.ystd, a.c0);
      const unsigned int word = __ballot_sync(0xffffffffu, valid && (y > a.tau));
      if (lane == g) mine = word;
      // rows only grow for a given lane, so a tie never replaces the earlier (smaller) index
      if (valid && y > st.bv[q]) { st.bv[q] = y; st.bi[q] = row; }
    }
    st.cnt[q] += __popc(mine);  // per lane (lanes 0..3 hold the four words); summed over the lanes at the flush
    if (q < nvalid) {
      if (wr) *mp = mine;
      pool |= mine;
    }
    mp += a.mask_ld;
  }
  if (a.pooled != nullptr && lane < TM / 32 && pool != 0u) atomicOr(a.pooled + (row0 >> 5) + lane, pool);
}

template <int NW>
__device__ __forceinline__ void epilogue_levelset_reg(const KArgs& a, const float* sOut, LsRegState<NW>& st,
                                                      int row0, int s0, int tid) {
  if ((long long)row0 + TM <= a.N) epilogue_levelset_reg_impl<NW, true>(a, sOut, st, row0, s0, tid);
  else epilogue_levelset_reg_impl<NW, false>(a, sOut, st, row0, s0, tid);
}

template <int NW>
__device__ __forceinline__ void epilogue_levelset_reg_flush(const KArgs& a, LsRegState<NW>& st, int s0, int tid) {
  constexpr int SPW = TN / NW;
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int q = 0; q < SPW; ++q) {
    float bv = st.bv[q];
    int bi = st.bi[q];
    unsigned int cnt = st.cnt[q];  // lanes 0..3 counted one row group each
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      const size_t o = (size_t)blockIdx.x * a.lay.Sp + s0 + warp * SPW + q;
      a.p_cnt[o] = cnt;
      a.p_max[o] = bv;
      a.p_arg[o] = bi;
    }
  }
}

// ---------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------
template <int DP, int EPI, bool BF>
__global__ void __launch_bounds__(TC_THREADS, 1) k_shared_tensor(const __grid_constant__ KArgs a) {
  using C = TcCfg<BF, DP == 0>;
  // no-swizzle UMMA operands only need 16-byte alignment; plain array arithmetic keeps the
  // shared address space visible to the compiler (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(128) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  const Layout& lay = a.lay;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TcSmem sl = tc_smem_layout(lay, EPI, a.k);
  uint8_t* sB = smem + sl.b;
  float* sFT = reinterpret_cast<float*>(smem + sl.ft);
  constexpr bool FT3 = PSBAX_TC_FT3 && DP == 2;
  float* sFT3 = reinterpret_cast<float*>(smem + sl.ft3);
  (void)sFT3;
  float* sOut = reinterpret_cast<float*>(smem + sl.out);
  float* sState = reinterpret_cast<float*>(smem + sl.state);
  float* sIl = reinterpret_cast<float*>(smem + sl.il);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sl.bars);
  uint64_t* b_full = bars;             // [8]
  uint64_t* b_empty = bars + 8;        // [8]
  uint64_t* a_full = bars + 16;        // [8]: [stage] or [stage][k-step]
  uint64_t* a_empty = bars + 24;       // [8]
  uint64_t* d_full = bars + 32;        // [2]
  uint64_t* d_empty = bars + 34;       // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 36);

  constexpr int TKB = C::TKB;
  constexpr bool GEN = (DP == 0);           // generic d: x in shared memory, table blocks streamed
  constexpr int TC_FPW = TKB / (TC_NPROD / 4);  // features per producer warp per block
  // feature blocks: nkb_r blocks of TKB rows (RFF / linear rows; every block in the 3xTF32 engine),
  // then (BF16x3 engine) nkb_k exact-kernel blocks of TKB / 2 rows contracted in 3xTF32
  const int nkb_r = lay.nkb_r;
  const int nkb = lay.nkb_r + lay.nkb_k;
  constexpr int KF = TKB / 2;               // rows of an exact-kernel block
  constexpr int OSC = (DP + 1 + 3) / 4 * 4; // row stride of the table for this DP (compile time)
  const int OS = GEN ? lay.OS : OSC;
  constexpr int NSB = C::NSB;
  const uint32_t STAGE = GEN ? sl.stage : (uint32_t)C::BBYTES;  // compile-time for the register-x kernels
  float* sX = reinterpret_cast<float*>(smem + sl.x);
  const int s0 = blockIdx.y * TN;
  const int npass = (a.ntiles + TC_TILES - 1) / TC_TILES;
  int t0, t1;  // this CTA's contiguous range of passes (pairs of row tiles)
  tc_tile_range(npass, gridDim.x, blockIdx.x, &t0, &t1);
  const EpiSmem e = epi_carve(sState, EPI, a.k);

  // ---- one-time setup ----
  {
    if constexpr (!GEN) {
      const float4* src = reinterpret_cast<const float4*>(a.pack + lay.off_ft);
      for (int i = tid; i < lay.Kp * OS / 4; i += TC_THREADS) reinterpret_cast<float4*>(sFT)[i] = __ldg(src + i);
    }
    for (int i = tid; i < round_up(lay.dp4, 4); i += TC_THREADS) sIl[i] = __ldg(a.pack + lay.off_invls + i);
    if constexpr (FT3) {
      const float* src = a.pack + lay.off_ft;
      for (int i = tid; i < lay.Kp * 3; i += TC_THREADS) sFT3[i] = __ldg(src + (i / 3) * OS + (i % 3));
    }
  }
  if (tid == 0) {
    // generic d: the producers read the streamed table block, so they release the stage too
    for (int i = 0; i < NSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], GEN ? 1 + TC_NPROD : 1); }
    for (int i = 0; i < C::NSA; ++i) { mbar_init(&a_full[i], TC_NPROD); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], TC_NEPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_WARP_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_NEPI) epi_init(e, EPI, a.k, tid - TC_EPI_WARP0 * 32, TC_NEPI * 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (PSBAX_EXP_FREERUN && !(warp >= TC_PROD_WARP0 && warp < TC_PROD_WARP0 + TC_NPROD)) {
  } else if (warp == TC_WARP_LOAD) {
    // ================= B loader =================
    const uint8_t* gB = reinterpret_cast<const uint8_t*>(a.pack + lay.off_tc) +
                        (size_t)blockIdx.y * nkb * C::BBYTES;
    const uint8_t* gFT = reinterpret_cast<const uint8_t*>(a.pack + lay.off_ft);
    const uint32_t ftbytes = (uint32_t)TKB * OS * 4;
    uint32_t it = 0;
    for (int pass = t0; pass < t1; ++pass) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t bs = it % NSB, ph = (it / NSB) & 1;
        if constexpr (PSBAX_TC_SLEEP_NS > 0) mbar_wait_hybrid_nanosleep<PSBAX_TC_SLEEP_NS / 4>(&b_empty[bs], ph ^ 1, lane);
        else mbar_wait_hybrid(&b_empty[bs], ph ^ 1, lane);
        if (elect_one()) {
          // table rows of the block (generic d): TKB rows from kb * TKB, or the KF rows of an exact-kernel block
          const bool kblk = BF && kb >= nkb_r;
          const uint32_t fb = kblk ? ftbytes / 2 : ftbytes;
          const size_t frow = kblk ? (size_t)lay.kk0 + (size_t)(kb - nkb_r) * KF : (size_t)kb * TKB;
#if PSBAX_EXP_NOCOPY  // what-if build (wrong results): the operand stream is not read at all (1), or only its
          // first NSB blocks, which then stay in the ring as realistic non-zero operands (2)
          if (PSBAX_EXP_NOCOPY == 2 && it < (uint32_t)NSB) {
            mbar_arrive_expect_tx(&b_full[bs], C::BBYTES + (GEN ? fb : 0u));
            bulk_g2s(sB + bs * STAGE, gB + (size_t)kb * C::BBYTES, C::BBYTES, &b_full[bs]);
          } else {
            mbar_arrive_expect_tx(&b_full[bs], (GEN ? fb : 0u));
          }
#else
          mbar_arrive_expect_tx(&b_full[bs], C::BBYTES + (GEN ? fb : 0u));
          bulk_g2s(sB + bs * STAGE, gB + (size_t)kb * C::BBYTES, C::BBYTES, &b_full[bs]);
#endif
          if constexpr (GEN)
            bulk_g2s(sB + bs * STAGE + C::BBYTES, gFT + frow * OS * 4, fb, &b_full[bs]);
        }
        __syncwarp();
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ================= MMA issuer =================
    // All lanes poll (converged), one elected lane issues (straight UTC*MMA runs).
    // tcgen05.mma issue blocks while the tensor pipe's short queue is full, so the readiness of
    // the NEXT block is probed in the middle of the current one: its latency is then covered by
    // queued MMAs instead of draining the pipe at the block boundary.
    constexpr uint32_t IDESC = BF ? umma_idesc_bf16(TN) : umma_idesc_tf32(TN);
    uint32_t it = 0, tl = 0;
    bool ready = false;
    for (int pass = t0; pass < t1; ++pass, ++tl) {
      const uint32_t buf = tl & 1, dph = (tl >> 1) & 1;
      mbar_wait(&d_empty[buf], dph ^ 1);  // all lanes poll: the single issuer warp stays converged
      __syncwarp();
      tc_fence_after();
      const uint32_t d_addr = tmem + TC_COL_D + buf * (TC_TILES * TN);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t bs = it % NSB;
        const uint32_t as = it % C::NSA, aph = (it / C::NSA) & 1;
        TC_TRACE(0, it, 0);
        const uint32_t a_base = tmem + TC_COL_A + as * (TC_TILES * C::ACT);
        const uint32_t bbase = smem_u32(sB + bs * STAGE);
        if (!ready) {
          mbar_wait(&b_full[bs], (it / NSB) & 1);
          mbar_wait(&a_full[as], aph);
          __syncwarp();
        }
        TC_TRACE(0, it, 1);
        tc_fence_after();
        // Issue order: k-step major, and inside a k-step the two row tiles alternate MMA by MMA, so that
        // consecutive MMAs accumulate into DIFFERENT tensor-memory tiles.  Twelve MMAs in a row into one
        // accumulator (the round-1 order: tile 0's whole block, then tile 1's) ran at about half the tensor
        // pipe's rate: same-box A/B 8.23 -> 7.84 ms at C2.  Per accumulator the order of additions is unchanged.
        if (elect_one()) {
          const uint32_t d0 = d_addr, d1 = d_addr + TN;
          if (BF && kb >= nkb_r) {
            // exact-kernel block: KF features as TF32 hi | lo (A columns [0, KF) | [KF, 2 KF)), K = 8 per MMA
            constexpr uint32_t IDESC_K = umma_idesc_tf32(TN);
            constexpr uint32_t SBO_K = KF * 4 * 8, SPLIT2_K = 8 * SBO_K;
#pragma unroll
            for (int ks = 0; ks < KF / 8; ++ks) {
              const uint64_t b1 = umma_desc_kmajor(bbase + ks * 256, 128, SBO_K);
              const uint64_t b2 = umma_desc_kmajor(bbase + SPLIT2_K + ks * 256, 128, SBO_K);
              const uint32_t a10 = a_base + ks * 8, a20 = a10 + C::ACT / 2;
              const uint32_t a11 = a10 + C::ACT, a21 = a11 + C::ACT / 2;
              tc_mma_ts(d0, a10, b1, IDESC_K, (kb | ks) ? 1u : 0u);
              tc_mma_ts(d1, a11, b1, IDESC_K, (kb | ks) ? 1u : 0u);
              tc_mma_ts(d0, a10, b2, IDESC_K, 1u);
              tc_mma_ts(d1, a11, b2, IDESC_K, 1u);
              tc_mma_ts(d0, a20, b1, IDESC_K, 1u);
              tc_mma_ts(d1, a21, b1, IDESC_K, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
              const uint64_t b1 = umma_desc_kmajor(bbase + ks * C::B_KSTEP, 128, C::B_SBO);
              const uint64_t b2 = umma_desc_kmajor(bbase + C::B_SPLIT2 + ks * C::B_KSTEP, 128, C::B_SBO);
              const uint32_t a10 = a_base + ks * C::A_KSTEP, a20 = a10 + C::A_SPLIT2;
              const uint32_t a11 = a10 + C::ACT, a21 = a11 + C::A_SPLIT2;
              if constexpr (BF) {
#if !PSBAX_EXP_NOMMA  // what-if build (wrong results): the issuer only commits
                tc_mma_ts_bf16(d0, a10, b1, IDESC, (kb | ks) ? 1u : 0u);
                tc_mma_ts_bf16(d1, a11, b1, IDESC, (kb | ks) ? 1u : 0u);
                tc_mma_ts_bf16(d0, a10, b2, IDESC, 1u);
                tc_mma_ts_bf16(d1, a11, b2, IDESC, 1u);
                tc_mma_ts_bf16(d0, a20, b1, IDESC, 1u);
                tc_mma_ts_bf16(d1, a21, b1, IDESC, 1u);
#endif
              } else {
                tc_mma_ts(d0, a10, b1, IDESC, (kb | ks) ? 1u : 0u);
                tc_mma_ts(d1, a11, b1, IDESC, (kb | ks) ? 1u : 0u);
                tc_mma_ts(d0, a10, b2, IDESC, 1u);
                tc_mma_ts(d1, a11, b2, IDESC, 1u);
                tc_mma_ts(d0, a20, b1, IDESC, 1u);
                tc_mma_ts(d1, a21, b1, IDESC, 1u);
              }
            }
          }
          // release the stages (and publish the accumulators)
          tc_commit(&b_empty[bs]);
          tc_commit(&a_empty[as]);
          if (kb == nkb - 1) tc_commit(&d_full[buf]);
        }
        __syncwarp();
        {  // probe block it + 1 while this block's MMAs are queued
          const uint32_t nit = it + 1;
          const bool ok = mbar_test(&a_full[nit % C::NSA], (nit / C::NSA) & 1) &&
                          mbar_test(&b_full[nit % NSB], (nit / NSB) & 1);
          ready = __all_sync(0xffffffffu, ok) != 0;
        }
        TC_TRACE(0, it, 2);
      }
    }
  } else if (warp >= TC_PROD_WARP0 && warp < TC_PROD_WARP0 + TC_NPROD) {
    // ================= producers =================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int h = (warp - TC_PROD_WARP0) >> 2;    // which TC_FPW features of every block: 0 .. TC_NPROD / 4 - 1
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
#if defined(PSBAX_TC_TRACE) && PSBAX_TC_TRACE == 2
    const bool tracer = (q == 0 && lane == 0);
    const int trole = h;
    (void)trole;
#else
    const bool tracer = (warp == TC_PROD_WARP0 && lane == 0);
#endif
    (void)tracer;
    uint32_t it = 0;
    auto load_rows = [&](int pass, uint64_t (&xr)[GEN ? 1 : DP]) {
      const long long row0 = (long long)pass * TC_TILES * TM + q * 32 + lane, row1 = row0 + TM;
#pragma unroll
      for (int i = 0; i < (GEN ? 0 : DP); ++i) {
        const float x0 = (row0 < a.N && i < lay.d) ? load_x(a, row0, i) : 0.f;
        const float x1 = (row1 < a.N && i < lay.d) ? load_x(a, row1, i) : 0.f;
        xr[i] = pack2(x0, x1);
      }
    };
    for (int pass = t0; pass < t1; ++pass) {
      static_assert(TC_TILES == 2, "the packed-FMA producers pair the two row tiles of a pass");
      uint64_t xp[GEN ? 1 : DP];  // (x of tile 0, x of tile 1) per dimension
      if constexpr (GEN) {
        // stage the pass's 256 rows transposed as (tile 0, tile 1) pairs: sX[i][row][tile], so that one
        // LDS.64 feeds a packed FFMA2 over the thread's two rows; producer-only named barrier
        constexpr int PT = TC_NPROD * 32;
        const int ptid = tid - TC_PROD_WARP0 * 32;
        asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");  // everyone done with the previous tile
        const long long base = (long long)pass * TC_TILES * TM * lay.d;
        const long long lim = a.N * lay.d;
        for (int idx = ptid; idx < TC_TILES * TM * lay.d; idx += PT) {
          const int rr = idx / lay.d, i = idx - rr * lay.d;  // rr in [0, 256)
          sX[(i * TM + (rr & 127)) * TC_TILES + (rr >> 7)] =
              (base + idx < lim) ? load_x(a, (long long)pass * TC_TILES * TM + rr, i) : 0.f;
        }
        for (int idx = ptid; idx < TC_TILES * TM * (lay.dp4 - lay.d); idx += PT)
          sX[lay.d * TM * TC_TILES + idx] = 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");
      } else {
        load_rows(pass, xp);
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t as = it % C::NSA, aph = (it / C::NSA) & 1;
        if (tracer) TC_TRACE_P(it, 0);
        const uint32_t bs = it % NSB;
        const bool kblk = BF && kb >= nkb_r;
        constexpr bool a_ready = PSBAX_EXP_FREERUN || PSBAX_EXP_NOAEMPTY;  // what-if builds never wait for the stage
        const int kbase0 = kb * TKB + h * TC_FPW;
        constexpr int NG = TC_FPW / 8;
        // EARLY: wait for the stage after the first 8-feature group and store every group at once
        // (lower register pressure: large DP / generic d); otherwise compute all groups before the
        // wait, so that its latency hides behind the math (measured better for small d)
        constexpr bool EARLY = GEN || DP >= 12;
        uint32_t s1[NG][TC_TILES][C::WCOLS], s2[NG][TC_TILES][C::WCOLS];
        if (kblk) {
          // ---- exact-kernel block of the BF16x3 engine: KF features, TF32 hi | lo operands ----
          // x in registers: KF = 32, 8 features per warp; generic d: KF = 16, 4 features per warp.
          const int kbk = kb - nkb_r;
          constexpr int KW = GEN ? 4 : 8;  // features per warp = tensor-memory columns per operand half
          uint32_t khi[TC_TILES][KW], klo[TC_TILES][KW];
          {
            float r2[TC_TILES][KW];
            if constexpr (GEN) {
              mbar_wait_hybrid(&b_full[bs], (it / NSB) & 1, lane);  // the block's table rows sit behind the B image
              const float* ft = reinterpret_cast<const float*>(sB + bs * STAGE + C::BBYTES) + h * KW * OS;
              uint64_t acc2[KW];
#pragma unroll
              for (int j = 0; j < KW; ++j) acc2[j] = 0ull;
              const uint64_t* xr = reinterpret_cast<const uint64_t*>(sX) + q * 32 + lane;
              for (int i0 = 0; i0 < lay.dp4; i0 += 4) {
                uint64_t xa[4];
                float il4[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { xa[c] = xr[(i0 + c) * TM]; il4[c] = sIl[i0 + c]; }
#pragma unroll
                for (int j = 0; j < KW; ++j) {
                  const float4 w4 = *reinterpret_cast<const float4*>(ft + j * OS + i0);
                  const float wv4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const uint64_t df = ffma2(xa[c], pack2(il4[c], il4[c]), pack2(wv4[c], wv4[c]));
                    acc2[j] = ffma2(df, df, acc2[j]);
                  }
                }
              }
#pragma unroll
              for (int j = 0; j < KW; ++j) unpack2(acc2[j], r2[0][j], r2[1][j]);
              __syncwarp();
            } else {
              const float* ft = sFT + (lay.kk0 + kbk * KF + h * KW) * OSC;
#pragma unroll
              for (int j = 0; j < KW; ++j) {
                float w[OSC];
                load_row<OSC>(ft + j * OSC, w);
                uint64_t acc = pack2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < DP; ++i) {
                  const uint64_t df = ffma2(xp[i], pack2(sIl[i], sIl[i]), pack2(w[i], w[i]));
                  acc = ffma2(df, df, acc);
                }
                unpack2(acc, r2[0][j], r2[1][j]);
              }
            }
#pragma unroll
            for (int t = 0; t < TC_TILES; ++t) {
              float kv[KW];
              kernel_profile_n<KW>(r2[t], kv, lay.kernel);
#pragma unroll
              for (int j = 0; j < KW; ++j) {
                const uint32_t hb = (__float_as_uint(kv[j]) + 0x1000u) & 0xffffe000u;  // round to TF32: |lo| <= 2^-11 |x|, unbiased
                khi[t][j] = hb;
                klo[t][j] = __float_as_uint(kv[j] - __uint_as_float(hb));
              }
            }
          }
          if (!a_ready) mbar_wait_hybrid_sleep<PSBAX_AEMPTY_HINT>(&a_empty[as], aph ^ 1, lane);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < TC_TILES; ++t) {
            const uint32_t col = TC_COL_A + (as * TC_TILES + t) * C::ACT + h * KW;
            if constexpr (GEN) {
              tc_st4(lane_addr + col, khi[t]);
              tc_st4(lane_addr + col + C::ACT / 2, klo[t]);
            } else {
              tc_st8(lane_addr + col, khi[t]);
              tc_st8(lane_addr + col + C::ACT / 2, klo[t]);
            }
          }
        } else {
        if constexpr (GEN) {  // the table block of this K-block arrives with the B image
          mbar_wait_hybrid(&b_full[bs], (it / NSB) & 1, lane);
        }
        const float* ftb = GEN ? reinterpret_cast<const float*>(sB + bs * STAGE + C::BBYTES) + h * TC_FPW * OS
                               : sFT + kbase0 * OS;
        // PURE: every feature of this warp's slice of the block is a cos feature (the common case):
        // the same body without the per-group section tests, selects and the kernel-profile branch
        auto group = [&](auto pure_tag, const int g) {
          constexpr bool PURE = decltype(pure_tag)::value;
          // 8 features for both tiles' rows: each table row is fetched once and used twice.
          // The RFF / kernel-feature branch is warp-uniform (sections are multiples of 8).
          const int kbase = kbase0 + g * 8;
          const float* ft = ftb + g * 8 * OS;
          const bool lin = PURE ? false : kbase >= lay.Lcos;  // linear-kernel rows: identity activation (warp-uniform)
          float v[TC_TILES][8];
          // BF16x3 engine: the rows between the RFF section and the first exact-kernel block are
          // padding (zero weights): nothing to evaluate
          const bool pad = BF && !PURE && kbase >= lay.Lp_ft;
          if (pad) {
#pragma unroll
            for (int t = 0; t < TC_TILES; ++t)
#pragma unroll
              for (int j = 0; j < 8; ++j) v[t][j] = 0.f;
          } else if constexpr (GEN) {
            // x pairs from shared memory, 4 dimensions at a time, reused by the 8 features; packed FP32
            // over the thread's two rows with the table entry (and inv_ls) as scalar-broadcast operands
            const bool rff = kbase < lay.Lp_ft;
            uint64_t acc2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float b0 = rff ? ft[j * OS + lay.dp4] : 0.f;
              acc2[j] = pack2(b0, b0);
            }
            const uint64_t* xr = reinterpret_cast<const uint64_t*>(sX) + q * 32 + lane;
            for (int i0 = 0; i0 < lay.dp4; i0 += 4) {
              uint64_t xa[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) xa[c] = xr[(i0 + c) * TM];
              if (rff) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 w4 = *reinterpret_cast<const float4*>(ft + j * OS + i0);
                  const float wv4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                  for (int c = 0; c < 4; ++c) acc2[j] = ffma2(xa[c], pack2(wv4[c], wv4[c]), acc2[j]);
                }
              } else {
                float il4[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) il4[c] = sIl[i0 + c];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 w4 = *reinterpret_cast<const float4*>(ft + j * OS + i0);
                  const float wv4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const uint64_t df = ffma2(xa[c], pack2(il4[c], il4[c]), pack2(wv4[c], wv4[c]));
                    acc2[j] = ffma2(df, df, acc2[j]);
                  }
                }
              }
            }
            float acc[TC_TILES][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) unpack2(acc2[j], acc[0][j], acc[1][j]);
            if (rff) {
#pragma unroll
              for (int t = 0; t < TC_TILES; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[t][j] = lin ? acc[t][j] : __cosf(acc[t][j]);
            } else {
#pragma unroll
              for (int t = 0; t < TC_TILES; ++t) kernel_profile8(acc[t], v[t], lay.kernel);
            }
          } else if (FT3 && PURE) {
            // dense rows: 24 floats = six 16-byte loads for the group's eight features
            float f[24];
            const float4* t4 = reinterpret_cast<const float4*>(sFT3 + kbase * 3);
#pragma unroll
            for (int u4 = 0; u4 < 6; ++u4) {
              const float4 qv = t4[u4];
              f[4 * u4] = qv.x; f[4 * u4 + 1] = qv.y; f[4 * u4 + 2] = qv.z; f[4 * u4 + 3] = qv.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint64_t arg = pack2(f[3 * j + 2], f[3 * j + 2]);
              arg = ffma2(xp[0], pack2(f[3 * j], f[3 * j]), arg);
              arg = ffma2(xp[DP - 1], pack2(f[3 * j + 1], f[3 * j + 1]), arg);
              float a0, a1;
              unpack2(arg, a0, a1);
              v[0][j] = __cosf(a0);
              v[1][j] = __cosf(a1);
            }
          } else if (PURE || kbase < lay.Lp_ft) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float w[OSC];
#if PSBAX_EXP_NOLDS  // what-if build (wrong results): table entries made up in registers, no shared-memory loads
#pragma unroll
              for (int i = 0; i < OSC; ++i) w[i] = __uint_as_float(0x3f000000u + (uint32_t)(kbase + j) * 64u + i * 7u);
#else
              load_row<OSC>(ft + j * OSC, w);
#endif
              uint64_t arg = pack2(w[DP], w[DP]);
#pragma unroll
              for (int i = 0; i < DP; ++i) arg = ffma2(xp[i], pack2(w[i], w[i]), arg);
              float a0, a1;
              unpack2(arg, a0, a1);
#if PSBAX_EXP_NOCOS  // what-if build (wrong results): no MUFU
              v[0][j] = a0 * 0.37f;
              v[1][j] = a1 * 0.37f;
#else
              v[0][j] = lin ? a0 : __cosf(a0);
              v[1][j] = lin ? a1 : __cosf(a1);
#endif
            }
          } else {
            float r2[TC_TILES][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float w[OSC];
              load_row<OSC>(ft + j * OSC, w);
              uint64_t acc = pack2(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < DP; ++i) {
                const uint64_t df = ffma2(xp[i], pack2(sIl[i], sIl[i]), pack2(w[i], w[i]));
                acc = ffma2(df, df, acc);
              }
              unpack2(acc, r2[0][j], r2[1][j]);
            }
#pragma unroll
            for (int t = 0; t < TC_TILES; ++t) kernel_profile8(r2[t], v[t], lay.kernel);
          }
          // split into the two MMA operands
          if constexpr (BF && !GEN) {
            // The values of one feature for the thread's two rows already sit in an aligned register
            // pair (packed projection), so the residual v - bf16(v) is formed per FEATURE over the
            // (tile 0, tile 1) pair; pairing two features of one row instead cost one MOV per value.
            static_assert(TC_TILES == 2, "row-pair split");
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // column c packs features 2c (low half) and 2c + 1 (high half)
              const uint32_t p1a = cvt_bf16x2(v[0][2 * c + 1], v[0][2 * c]);
              const uint32_t p1b = cvt_bf16x2(v[1][2 * c + 1], v[1][2 * c]);
#if PSBAX_SPLIT_FHFMA
              float a0, a1, b0, b1;  // residuals: a = feature 2c, b = feature 2c + 1; 0 / 1 = tile
              bf16_residual2(p1a, v[0][2 * c], v[0][2 * c + 1], a0, b0);
              bf16_residual2(p1b, v[1][2 * c], v[1][2 * c + 1], a1, b1);
#else
              const uint64_t ra = ffma2(pack2(__uint_as_float(p1a << 16), __uint_as_float(p1b << 16)),
                                        pack2(-1.0f, -1.0f), pack2(v[0][2 * c], v[1][2 * c]));
              const uint64_t rb = ffma2(pack2(__uint_as_float(p1a & 0xffff0000u), __uint_as_float(p1b & 0xffff0000u)),
                                        pack2(-1.0f, -1.0f), pack2(v[0][2 * c + 1], v[1][2 * c + 1]));
              float a0, a1, b0, b1;
              unpack2(ra, a0, a1);
              unpack2(rb, b0, b1);
#endif
              s1[g][0][c] = p1a;
              s1[g][1][c] = p1b;
              s2[g][0][c] = cvt_bf16x2(b0, a0);
              s2[g][1][c] = cvt_bf16x2(b1, a1);
            }
          } else {
#pragma unroll
          for (int t = 0; t < TC_TILES; ++t) {
            if constexpr (BF) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {  // column c packs features 2c (low half) and 2c + 1 (high half)
                const uint32_t p1 = cvt_bf16x2(v[t][2 * c + 1], v[t][2 * c]);
                float r0, r1;
#if PSBAX_SPLIT_FHFMA
                bf16_residual2(p1, v[t][2 * c], v[t][2 * c + 1], r0, r1);
#else
                // residual pair (v - bf16(v)) as one packed FMA: (-1) * x1 + v
                const uint64_t rr = ffma2(pack2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u)),
                                          pack2(-1.0f, -1.0f), pack2(v[t][2 * c], v[t][2 * c + 1]));
                unpack2(rr, r0, r1);
#endif
                s1[g][t][c] = p1;
                s2[g][t][c] = cvt_bf16x2(r1, r0);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t hb = (__float_as_uint(v[t][j]) + 0x1000u) & 0xffffe000u;  // round to TF32
                s1[g][t][j] = hb;
                s2[g][t][j] = __float_as_uint(v[t][j] - __uint_as_float(hb));
              }
            }
          }
          }
          if (EARLY ? (g == 0) : (g == NG - 1)) {
            if (tracer) TC_TRACE_P(it, 1);
            if (!a_ready) mbar_wait_hybrid_sleep<PSBAX_AEMPTY_HINT>(&a_empty[as], aph ^ 1, lane);
            if (tracer) TC_TRACE_P(it, 2);
            tc_fence_after();
          }
          if constexpr (!EARLY && BF && NG == 2) {
            // a warp's 16 features are one k-step: both splits of both groups leave as one 16-column store per tile
            if (g == NG - 1) {
#pragma unroll
              for (int t = 0; t < TC_TILES; ++t) {
                static_assert(C::INTERLEAVED && NG * C::WCOLS == C::KCOLS, "one k-step per producer warp");
                const uint32_t col = TC_COL_A + (as * TC_TILES + t) * C::ACT + h * C::A_KSTEP;
                const uint32_t u[16] = {s1[0][t][0], s1[0][t][1], s1[0][t][2], s1[0][t][3],
                                        s1[1][t][0], s1[1][t][1], s1[1][t][2], s1[1][t][3],
                                        s2[0][t][0], s2[0][t][1], s2[0][t][2], s2[0][t][3],
                                        s2[1][t][0], s2[1][t][1], s2[1][t][2], s2[1][t][3]};
#if PSBAX_EXP_NOSTTM  // what-if build (wrong results): values consumed by a dummy, no tensor-memory store
                uint32_t x = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) x ^= u[i];
                if (x == 0x12345678u) sIl[0] = 1.f;
#else
                tc_st16(lane_addr + col, u);
#endif
              }
            }
          } else if (EARLY || g == NG - 1) {
#pragma unroll
            for (int gg = EARLY ? g : 0; gg <= g; ++gg)
#pragma unroll
              for (int t = 0; t < TC_TILES; ++t) {
                const uint32_t col = TC_COL_A + (as * TC_TILES + t) * C::ACT +
                                     (C::INTERLEAVED ? h * C::A_KSTEP + gg * C::WCOLS : (h * NG + gg) * C::WCOLS);
                if constexpr (BF) {
                  tc_st4(lane_addr + col, s1[gg][t]);
                  tc_st4(lane_addr + col + C::A_SPLIT2, s2[gg][t]);
                } else {
                  tc_st8(lane_addr + col, s1[gg][t]);
                  tc_st8(lane_addr + col + C::ACT / 2, s2[gg][t]);
                }
              }
          }
        };
        if (!GEN && kbase0 + TC_FPW <= lay.Lcos) {
#pragma unroll
          for (int g = 0; g < NG; ++g) group(std::true_type{}, g);
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) group(std::false_type{}, g);
        }
        }  // !kblk
        if constexpr (GEN) {  // done reading the streamed table block
          __syncwarp();
          if (lane == 0) mbar_arrive(&b_empty[bs]);
        }
        tc_wait_st();
        if (tracer) TC_TRACE_P(it, 3);
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && !PSBAX_EXP_FREERUN) mbar_arrive(&a_full[as]);
        if (tracer) TC_TRACE_P(it, 4);
      }
    }
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_NEPI) {
    // ================= epilogue =================
    const int q = warp & 3;
    const int etid = tid - TC_EPI_WARP0 * 32;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    LsRegState<TC_NEPI> ls;
    ls.init();
    uint32_t tl = 0;
    for (int pass = t0; pass < t1; ++pass, ++tl) {
      const uint32_t buf = tl & 1, dph = (tl >> 1) & 1;
      if (warp == TC_EPI_WARP0 && lane == 0) TC_TRACE(2, tl, 0);
      if constexpr (PSBAX_TC_SLEEP_NS > 0) mbar_wait_hybrid_nanosleep<PSBAX_TC_SLEEP_NS>(&d_full[buf], dph, lane);
      else mbar_wait_hybrid_sleep<20000>(&d_full[buf], dph, lane);
      if (warp == TC_EPI_WARP0 && lane == 0) TC_TRACE(2, tl, 1);
      tc_fence_after();
      const int r = q * 32 + lane;
#pragma unroll 1
      for (int t = 0; t < TC_TILES; ++t) {
        const int tile = pass * TC_TILES + t;
        if (tile < a.ntiles) {
#pragma unroll 1
          for (int c = 0; c < TN / 16; ++c) {
            uint32_t u[16];
            tc_ld16(lane_addr + TC_COL_D + (buf * TC_TILES + t) * TN + c * 16, u);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) sOut[(c * 16 + j) * OUT_LD + r] = __uint_as_float(u[j]);
          }
        }
        if (t == TC_TILES - 1) {  // both tiles' accumulators have left tensor memory: release D[buf]
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&d_empty[buf]);
          if (warp == TC_EPI_WARP0 && lane == 0) TC_TRACE(2, tl, 2);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
        if (tile < a.ntiles && !PSBAX_EXP_NOEPI) {
          if constexpr (EPI == EPI_LEVELSET) epilogue_levelset_reg<TC_NEPI>(a, sOut, ls, tile * TM, s0, etid);
          else epilogue_tile<EPI, TC_NEPI>(a, sOut, e, tile * TM, s0, etid);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
      }
      if (warp == TC_EPI_WARP0 && lane == 0) TC_TRACE(2, tl, 3);
    }
    if constexpr (EPI == EPI_LEVELSET) epilogue_levelset_reg_flush<TC_NEPI>(a, ls, s0, etid);
    else epilogue_flush<EPI>(a, e, s0, etid, TC_NEPI * 32);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_ALLOC) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
}

template <int EPI, bool BF>
int tensor_launch_epi(const KArgs& a0, int sms, cudaStream_t st, char* err, size_t errlen) {
  KArgs a = a0;
  const Layout& l = a.lay;
  const int gx = tensor_parts(l, a.N, sms), gy = l.Sp / TN;
  const TcSmem sl = tc_smem_layout(l, EPI, a.k);
  if (sl.total > TC_SMEM_MAX) {
    snprintf(err, errlen, "tensor engine: shared memory %u B exceeds 227 KB (L + n or d too large)", sl.total);
    return PSBAX_E_UNSUPPORTED;
  }
  cudaError_t ce = cudaSuccess;
#define TC_LAUNCH(DPV)                                                                                  \
  do {                                                                                                  \
    ce = cudaFuncSetAttribute(k_shared_tensor<DPV, EPI, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                              (int)sl.total);                                                           \
    if (ce == cudaSuccess) k_shared_tensor<DPV, EPI, BF><<<dim3(gx, gy), TC_THREADS, sl.total, st>>>(a);  \
  } while (0)
  switch (l.tc_generic ? 0 : l.dp4) {
    case 2: TC_LAUNCH(2); break;
    case 4: TC_LAUNCH(4); break;
    case 8: TC_LAUNCH(8); break;
    case 12: TC_LAUNCH(12); break;
    case 16: TC_LAUNCH(16); break;
    default: TC_LAUNCH(0); break;  // generic d / large tables
  }
#undef TC_LAUNCH
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "tensor engine launch failed: %s", cudaGetErrorString(ce));
    return PSBAX_E_CUDA;
  }
  return PSBAX_OK;
}

// one translation unit per operand format instantiates the kernels (tensor_bf16.cu / tensor_tf32.cu)
int tensor_launch_bf16(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen);
int tensor_launch_tf32(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen);
template <bool BF>
int tensor_launch_fmt(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  if (epi == EPI_VALUES) return tensor_launch_epi<EPI_VALUES, BF>(a, sms, st, err, errlen);
  if (epi == EPI_LEVELSET) return tensor_launch_epi<EPI_LEVELSET, BF>(a, sms, st, err, errlen);
  return tensor_launch_epi<EPI_TOPK, BF>(a, sms, st, err, errlen);
}
inline int tensor_launch(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  return a.lay.mode == MODE_TENSOR_BF16 ? tensor_launch_bf16(epi, a, sms, st, err, errlen)
                                        : tensor_launch_tf32(epi, a, sms, st, err, errlen);
}

}  // namespace psbax
