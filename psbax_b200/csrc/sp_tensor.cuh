// tcgen05 engine: shared-basis sample paths as a 3xTF32 tensor-core contraction whose
// A operand (the RFF / kernel features) is generated on the fly and never leaves the SM.
//
// One persistent CTA per SM, 16 warps, warp-specialised:
//   warp 0      B loader   : bulk-async copies (cp.async.bulk, TMA unit) of 16 KB blocks of the
//                            hi/lo-split [W;V] operand, pre-laid-out in UMMA canonical K-major
//                            form by tensor_pack(), into a 4-stage shared-memory ring
//   warp 1      MMA issuer : one thread issues tcgen05.mma kind::tf32, A from TENSOR MEMORY,
//                            B from shared memory, fp32 accumulators in tensor memory
//   warp 2      TMEM allocator
//   warps 4-11  producers  : lane = candidate row (x in registers).  Per 32-feature block each
//                            thread evaluates 16 features (FFMA projection + MUFU cos, or the
//                            Matern/RBF profile), splits them into TF32 hi/lo and writes them
//                            straight into the A stage in tensor memory (tcgen05.st)
//   warps 12-15 epilogue   : tcgen05.ld the finished 128 x 64 tile, stage it in shared memory
//                            and run the same fused reduction as the SIMT kernels (level-set
//                            ballots / top-k filter / values dump) while the next tile's MMAs run
//
// 3xTF32:  f = Phi_hi*W_hi + Phi_hi*W_lo + Phi_lo*W_hi  (fp32 accumulate), issued as
//   MMA1: A = Phi_hi, B = [W_hi | W_lo]  (N = 128)  -> D[:, 0:128]
//   MMA2: A = Phi_lo, B =  W_hi          (N =  64)  -> D[:, 0:64]  (accumulate)
// and the epilogue adds D[:, s] + D[:, 64 + s].
//
// Tensor-memory columns (512 allocated): D0 [0,128) D1 [128,256) | A stage s: hi [256+64s, +32),
// lo [288+64s, +32).
#pragma once
#include "sp_common.cuh"
#include "sp_simt.cuh"

namespace psbax {

constexpr int TC_MAX_D = 16;
constexpr int TC_MAX_S = 1 << 20;  // any S: processed in chunks of 64 samples
constexpr int TC_THREADS = 512;
constexpr int TC_NSB = 4;                    // B stages
constexpr int TC_BBYTES = 128 * TK * 4;      // 16 KB per stage: 128 rows (hi|lo) x 32 k
constexpr int TC_PROD_WARP0 = 4, TC_NPROD = 8;
constexpr int TC_EPI_WARP0 = 12, TC_NEPI = 4;
constexpr int TC_COL_D = 0, TC_COL_A = 256;
constexpr uint32_t TC_TMEM_COLS = 512;

inline bool tensor_supported(const psbax_paths_t* p) {
  return p->G == 1 && p->S >= 16 && p->L > 0 && p->d <= TC_MAX_D;
}
// [chunks][nkb] images of 16 KB
inline size_t tensor_pack_floats(const Layout& l) {
  return (size_t)(l.Sp / TN) * (l.Kp / TK) * (TC_BBYTES / 4);
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
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
// instruction descriptor: TF32 x TF32 -> F32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct TcSmem {
  // offsets in bytes from the 1024-aligned dynamic smem base
  uint32_t b, ft, out, state, il, bars, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(const Layout& l, int epi, int k) {
  TcSmem s{};
  uint32_t o = 0;
  s.b = o;     o += TC_NSB * TC_BBYTES;
  s.ft = o;    o += (uint32_t)l.Kp * l.OS * 4;
  s.out = o;   o += TN * OUT_LD * 4;
  s.state = o; o += (uint32_t)round_up((int)epi_smem_floats(epi, k), 4) * 4;
  s.il = o;    o += (uint32_t)round_up(l.dp4, 4) * 4;
  o = (o + 15) / 16 * 16;
  s.bars = o;  o += 32 * 8;
  s.total = o + 1024;  // slack for the manual 1024-byte alignment
  return s;
}

// CTA c of gx handles the contiguous tile range [t0, t1)
__host__ __device__ inline void tc_tile_range(int ntiles, int gx, int c, int* t0, int* t1) {
  const int base = ntiles / gx, rem = ntiles % gx;
  *t0 = c * base + (c < rem ? c : rem);
  *t1 = *t0 + base + (c < rem ? 1 : 0);
}

inline int tensor_parts(const Layout& l, long long N, int sms) {
  const long long nt = (N + TM - 1) / TM;
  const int gy = l.Sp / TN;
  long long gx = (sms + gy - 1) / gy;
  if (gx > nt) gx = nt;
  if (gx < 1) gx = 1;
  return (int)gx;
}

// ---------------------------------------------------------------------------------
// operand pack: hi/lo TF32 split of [W;V] in UMMA canonical (no-swizzle, K-major) layout
//   image(cy, kb)[row r in 0..127][k in 0..31]:  r < 64: hi(wv[kb*32+k][cy*64+r]),
//                                                r >= 64: lo(wv[kb*32+k][cy*64+r-64])
//   float offset inside the image: (r/8)*256 + (k/4)*32 + (r%8)*4 + (k%4)
// ---------------------------------------------------------------------------------
__global__ void k_tensor_pack(Layout lay, float* __restrict__ pack) {
  const float* __restrict__ wv = pack + lay.off_wv;
  float* __restrict__ tc = pack + lay.off_tc;
  const int nkb = lay.Kp / TK;
  const size_t total = (size_t)(lay.Sp / TN) * nkb * 4096;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int within = (int)(idx & 4095);
    const size_t img = idx >> 12;
    const int kb = (int)(img % nkb), cy = (int)(img / nkb);
    const int rg = within >> 8, kc = (within >> 5) & 7, r8 = (within >> 2) & 7, k4 = within & 3;
    const int r = rg * 8 + r8, k = kc * 4 + k4;
    const float x = wv[(size_t)(kb * TK + k) * lay.Sp + cy * TN + (r & 63)];
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);  // round to TF32
    tc[idx] = (r < 64) ? hi : (x - hi);
  }
}

inline int tensor_pack(const Layout& l, float* pack, cudaStream_t st) {
  k_tensor_pack<<<296, 256, 0, st>>>(l, pack);
  return cudaGetLastError() == cudaSuccess ? PSBAX_OK : PSBAX_E_CUDA;
}

// ---------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------
template <int DP, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) k_shared_tensor(const __grid_constant__ KArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const Layout& lay = a.lay;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const TcSmem sl = tc_smem_layout(lay, EPI, a.k);
  uint8_t* sB = smem + sl.b;
  float* sFT = reinterpret_cast<float*>(smem + sl.ft);
  float* sOut = reinterpret_cast<float*>(smem + sl.out);
  float* sState = reinterpret_cast<float*>(smem + sl.state);
  float* sIl = reinterpret_cast<float*>(smem + sl.il);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sl.bars);
  uint64_t* b_full = bars;                 // [TC_NSB]
  uint64_t* b_empty = bars + TC_NSB;       // [TC_NSB]
  uint64_t* a_full = bars + 2 * TC_NSB;    // [2]
  uint64_t* a_empty = a_full + 2;          // [2]
  uint64_t* d_full = a_empty + 2;          // [2]
  uint64_t* d_empty = d_full + 2;          // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int nkb = lay.Kp / TK;
  const int OS = lay.OS;
  const int s0 = blockIdx.y * TN;
  int t0, t1;
  tc_tile_range(a.ntiles, gridDim.x, blockIdx.x, &t0, &t1);
  const EpiSmem e = epi_carve(sState, EPI, a.k);

  // ---- one-time setup ----
  {
    const float4* src = reinterpret_cast<const float4*>(a.pack + lay.off_ft);
    for (int i = tid; i < lay.Kp * OS / 4; i += TC_THREADS) reinterpret_cast<float4*>(sFT)[i] = __ldg(src + i);
    for (int i = tid; i < round_up(lay.dp4, 4); i += TC_THREADS) sIl[i] = __ldg(a.pack + lay.off_invls + i);
  }
  if (tid == 0) {
    for (int i = 0; i < TC_NSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], TC_NPROD);
      mbar_init(&a_empty[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], TC_NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= TC_EPI_WARP0) epi_init(e, EPI, a.k, tid - TC_EPI_WARP0 * 32, TC_NEPI * 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    // ================= B loader =================
    if (lane == 0) {
      const uint8_t* gB = reinterpret_cast<const uint8_t*>(a.pack + lay.off_tc) +
                          (size_t)blockIdx.y * nkb * TC_BBYTES;
      uint32_t it = 0;
      for (int tile = t0; tile < t1; ++tile) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t bs = it % TC_NSB, ph = (it / TC_NSB) & 1;
          mbar_wait(&b_empty[bs], ph ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], TC_BBYTES);
          bulk_g2s(sB + bs * TC_BBYTES, gB + (size_t)kb * TC_BBYTES, TC_BBYTES, &b_full[bs]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC_N128 = umma_idesc_tf32(128);
      constexpr uint32_t IDESC_N64 = umma_idesc_tf32(64);
      uint32_t it = 0, tl = 0;
      for (int tile = t0; tile < t1; ++tile, ++tl) {
        const uint32_t buf = tl & 1, dph = (tl >> 1) & 1;
        mbar_wait(&d_empty[buf], dph ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem + TC_COL_D + buf * 128;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t bs = it % TC_NSB, bph = (it / TC_NSB) & 1;
          const uint32_t as = it & 1, aph = (it >> 1) & 1;
          mbar_wait(&b_full[bs], bph);
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_hi = tmem + TC_COL_A + as * 64, a_lo = a_hi + 32;
          const uint32_t bbase = smem_u32(sB + bs * TC_BBYTES);
#pragma unroll
          for (int ks = 0; ks < TK / 8; ++ks) {
            const uint64_t bd = umma_desc_kmajor(bbase + ks * 256, 128, 1024);
            tc_mma_ts(d_addr, a_hi + ks * 8, bd, IDESC_N128, (kb | ks) ? 1u : 0u);
            tc_mma_ts(d_addr, a_lo + ks * 8, bd, IDESC_N64, 1u);
          }
          tc_commit(&b_empty[bs]);
          tc_commit(&a_empty[as]);
        }
        tc_commit(&d_full[buf]);
      }
    }
  } else if (warp >= TC_PROD_WARP0 && warp < TC_PROD_WARP0 + TC_NPROD) {
    // ================= producers =================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int h = (warp - TC_PROD_WARP0) >> 2;    // which 16 of the block's 32 features
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile) {
      const long long row = (long long)tile * TM + q * 32 + lane;
      float x[DP], xs[DP];
#pragma unroll
      for (int i = 0; i < DP; ++i) {
        x[i] = (row < a.N && i < lay.d) ? __ldg(a.X + row * lay.d + i) : 0.f;
        xs[i] = x[i] * sIl[i];
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        uint32_t hi[16], lo[16];
        const int kbase = kb * TK + h * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const bool rff = (kbase + g * 8) < lay.Lp_ft;  // warp-uniform, sections are multiples of 8
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* ft = sFT + (kbase + g * 8 + j) * OS;
            float v;
            if (rff) v = __cosf(rff_arg<DP>(ft, x));
            else v = kernel_profile(kern_r2<DP>(ft, xs), lay.kernel);
            const uint32_t hb = __float_as_uint(v) & 0xffffe000u;
            hi[g * 8 + j] = hb;
            lo[g * 8 + j] = __float_as_uint(v - __uint_as_float(hb));
          }
        }
        mbar_wait(&a_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t col = TC_COL_A + as * 64 + h * 16;
        tc_st16(lane_addr + col, hi);
        tc_st16(lane_addr + col + 32, lo);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
      }
    }
  } else if (warp >= TC_EPI_WARP0) {
    // ================= epilogue =================
    const int q = warp & 3;
    const int etid = tid - TC_EPI_WARP0 * 32;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t tl = 0;
    for (int tile = t0; tile < t1; ++tile, ++tl) {
      const uint32_t buf = tl & 1, dph = (tl >> 1) & 1;
      mbar_wait(&d_full[buf], dph);
      tc_fence_after();
      const int r = q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < TN / 16; ++c) {
        uint32_t u[16], w[16];
        tc_ld16(lane_addr + TC_COL_D + buf * 128 + c * 16, u);
        tc_ld16(lane_addr + TC_COL_D + buf * 128 + 64 + c * 16, w);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          sOut[(c * 16 + j) * OUT_LD + r] = __uint_as_float(u[j]) + __uint_as_float(w[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[buf]);
      asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
      epilogue_tile<EPI, TC_NEPI>(a, sOut, e, tile * TM, s0, etid);
      asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
    }
    epilogue_flush<EPI>(a, e, s0, etid, TC_NEPI * 32);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
}

template <int EPI>
int tensor_launch_epi(const KArgs& a0, int sms, cudaStream_t st, char* err, size_t errlen) {
  KArgs a = a0;
  const Layout& l = a.lay;
  const int gx = tensor_parts(l, a.N, sms), gy = l.Sp / TN;
  const TcSmem sl = tc_smem_layout(l, EPI, a.k);
  if (sl.total > 227 * 1024) {
    snprintf(err, errlen, "tensor engine: shared memory %u B exceeds 227 KB (L + n or d too large)", sl.total);
    return PSBAX_E_UNSUPPORTED;
  }
  cudaError_t ce = cudaSuccess;
#define TC_LAUNCH(DPV)                                                                              \
  do {                                                                                              \
    ce = cudaFuncSetAttribute(k_shared_tensor<DPV, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                              (int)sl.total);                                                       \
    if (ce == cudaSuccess) k_shared_tensor<DPV, EPI><<<dim3(gx, gy), TC_THREADS, sl.total, st>>>(a);  \
  } while (0)
  switch (l.dp4) {
    case 2: TC_LAUNCH(2); break;
    case 4: TC_LAUNCH(4); break;
    case 8: TC_LAUNCH(8); break;
    case 12: TC_LAUNCH(12); break;
    case 16: TC_LAUNCH(16); break;
    default:
      snprintf(err, errlen, "tensor engine: d=%d not supported", l.d);
      return PSBAX_E_UNSUPPORTED;
  }
#undef TC_LAUNCH
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "tensor engine launch failed: %s", cudaGetErrorString(ce));
    return PSBAX_E_CUDA;
  }
  return PSBAX_OK;
}

inline int tensor_launch(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  if (epi == EPI_VALUES) return tensor_launch_epi<EPI_VALUES>(a, sms, st, err, errlen);
  if (epi == EPI_LEVELSET) return tensor_launch_epi<EPI_LEVELSET>(a, sms, st, err, errlen);
  return tensor_launch_epi<EPI_TOPK>(a, sms, st, err, errlen);
}

}  // namespace psbax
