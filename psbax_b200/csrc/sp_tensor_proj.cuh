// tcgen05 engine, projection variant (5 <= d <= 32): the feature arguments omega_l . x + b_l are
// themselves computed by the tensor cores, so the producers only do MUFU + the operand split.
//
//   Z[256 rows x 32 features] = X_aug[256 x Kx] * Omega_aug^T[Kx x 32]      (3xTF32, fp32 accumulate)
//       X_aug = (x_0..x_{d-1}, 1, 0..)  hi/lo TF32, written ONCE PER PASS into tensor memory by the
//               producers (lane = row), A operand of the projection MMAs (TS mode); when d % 8 == 0
//               there is no spare column and the producers add the bias instead
//       Omega_aug = (omega_l | bias_l | 0..) hi/lo TF32 images streamed by the loader behind the
//               [W;V] image of the same 32-feature block, B operand
//   producers: tcgen05.ld their 8 arguments x 2 tiles from Z, cos (or identity for linear-kernel
//               rows; exact-kernel features still take the FFMA path from a shared-memory x tile),
//               BF16x3 split, tcgen05.st into the A stage
//   contraction and fused reductions as in k_shared_tensor (BF16x3, 32-feature blocks).
//
// Warps: 0-15 producers, 16-19 epilogue, 20 contraction issuer, 21 loader, 22 TMEM allocator,
// 23 projection issuer (it runs up to two blocks ahead of the producers: Z is double-buffered).
//
// Two shapes of a CTA pass (template parameter NCH = sample chunks of 64 per CTA):
//   NCH = 1 (S <= 64)  : TWO 128-row tiles x 64 samples.  Tensor-memory columns: D[tile] 0..127
//                        (single-buffered: both tiles are drained to two shared staging tiles before
//                        the reduction runs) | X[tile][split] 128..255 | Z[buf][tile] 256..383 |
//                        A[stage][tile] 384..511 (2 stages).
//   NCH = 2, 4 (S > 64): ONE 128-row tile x NCH * 64 samples: every feature of a candidate is
//                        generated once and contracted with all the chunks' [W;V] images (before: one
//                        CTA column per chunk, i.e. cos / kernel profiles evaluated S / 64 times).
//                        D[chunk] 0..NCH*64 | X[split] 64 | Z[buf] 64 | A[stage] 32 each, 4 stages.
#pragma once
#include "sp_tensor.cuh"

namespace psbax {

constexpr int TP_MIN_D = 5, TP_MAX_D = 32;
constexpr int TP_AUTO_MIN_D = 5;  // AUTO uses the projection variant from this d on (measured: wins from d = 5)
constexpr int TP_NSB_MAX = 8, TP_NSB_MIN = 3;  // operand ring depth: as deep as shared memory allows
constexpr int TP_WARP_PROJ = TC_WARP_MMA + 3;  // the otherwise idle fourth service warp
static_assert(TP_WARP_PROJ * 32 < TC_THREADS, "projection issuer warp");
constexpr int TP_TKB = 32;
// tensor-memory map of a pass shape
template <int NCH>
struct TpCols {
  static constexpr int TILES = NCH == 1 ? 2 : 1;
  static constexpr int D = 0;
  static constexpr int X = NCH * TILES * TN;          // behind the accumulators
  static constexpr int Z = X + TILES * 2 * 32;        // X: [tile][split] of 32 columns
  static constexpr int A = Z + 2 * TILES * 32;        // Z: [buf][tile] of 32 columns
  static constexpr int NSA_FIT = (512 - A) / (TILES * 32);
  static constexpr int NSA = NSA_FIT > 4 ? 4 : NSA_FIT;  // A ring depth (a power of two)
  static_assert(NSA == 2 || NSA == 4, "A ring depth");
};
constexpr int TP_WBYTES = 128 * TP_TKB * 2;  // bf16 [W;V] image of a block: 8 KB

// K extent of the projection.  The bias rides as an extra column of X_aug (= 1) when the padding
// to a multiple of 8 leaves room for it; for d % 8 == 0 the producers add it instead (keeps
// d = 32 inside the 32 tensor-memory columns per X split).
__host__ __device__ inline bool tp_bias_in_x(int d) { return (d % 8) != 0; }
__host__ __device__ inline int tp_kx(int d) { return tp_bias_in_x(d) ? round_up(d + 1, 8) : d; }
__host__ __device__ inline uint32_t tp_obytes(int d) { return 2u * 32u * (uint32_t)tp_kx(d) * 4u; }  // hi + lo image

struct TpSmem {
  // `b`: ring of `nsb` stages = nch [W;V] images | Omega_aug images | the block's rows of the feature
  // table (only when the problem has exact-kernel rows; copied only for blocks that contain them)
  uint32_t b, stage, nsb, ftk, x, out, state, state_stride, il, bias, seq, bars, total;
};
__host__ __device__ inline TpSmem tp_smem_layout(const Layout& l, int epi, int k, int nch = 1) {
  TpSmem s{};
  const uint32_t tiles = nch == 1 ? 2u : 1u;
  s.ftk = (uint32_t)nch * TP_WBYTES + tp_obytes(l.d);  // offset of the table block inside a stage
  s.stage = s.ftk + (l.np_ > 0 ? (uint32_t)(TP_TKB / 2) * l.OS * 4 : 0u);  // + the 16 table rows of an exact-kernel block
  s.state_stride = (uint32_t)round_up((int)epi_smem_floats(epi, k), 4) * 4;  // reduction state of one sample chunk
  // everything but the ring first; the ring takes what is left (a stage is released only after
  // its contraction completes and a bulk copy takes ~2000 cycles, so depth 4 starved the issuers)
  const uint32_t xbytes = l.n > 0 ? tiles * l.dp4 * TM * 4 : 0u;
  const uint32_t obytes = tiles * TN * OUT_LD * 4;  // staging: both tiles (NCH = 1) / one chunk at a time
  const uint32_t seqbytes = (uint32_t)round_up((l.nkb_r + l.nkb_k) * 2, 16);  // block order of a pass (uint16 per block)
  const uint32_t fixed = xbytes + obytes + (uint32_t)nch * s.state_stride + (uint32_t)round_up(l.dp4, 4) * 4 +
                         (tp_bias_in_x(l.d) ? 0u : (uint32_t)l.Kp * 4) + seqbytes + 16 + 40 * 8;
  uint32_t nsb = fixed < 227u * 1024u ? (227u * 1024u - fixed) / s.stage : 0u;
  s.nsb = nsb > (uint32_t)TP_NSB_MAX ? (uint32_t)TP_NSB_MAX : nsb;
  uint32_t o = 0;
  s.b = o;     o += s.nsb * s.stage;
  s.x = o;     o += xbytes;                                            // x tile for the kernel rows
  s.out = o;   o += obytes;
  s.state = o; o += (uint32_t)nch * s.state_stride;
  s.il = o;    o += (uint32_t)round_up(l.dp4, 4) * 4;
  s.bias = o;  o += tp_bias_in_x(l.d) ? 0u : (uint32_t)l.Kp * 4;       // per-feature bias (d % 8 == 0)
  o = (o + 15) / 16 * 16;
  s.seq = o;   o += seqbytes;
  s.bars = o;  o += 40 * 8;
  s.total = o;
  return s;
}
// shape rule for the projection variant (BF16x3 engine only)
__host__ __device__ inline bool tp_supported(const Layout& l) {
  if (TC_NPROD != 16) return false;  // the kernel assumes four producer warps per lane quarter
  if (l.mode != MODE_TENSOR_BF16 || l.d < TP_MIN_D || l.d > TP_MAX_D) return false;
  const TpSmem sl = tp_smem_layout(l, EPI_TOPK, KMAX);
  return sl.nsb >= (uint32_t)TP_NSB_MIN && sl.total <= TC_SMEM_MAX;
}
inline bool tensor_proj_supported(const psbax_paths_t* p) {
  if (!tensor_supported(p, true)) return false;
  const bool lin = p->kernel == PSBAX_KERNEL_LINEAR;
  Layout l{};
  l.mode = MODE_TENSOR_BF16;
  l.d = p->d;
  l.n = lin ? 0 : p->n;
  l.dp4 = (p->d <= 2) ? 2 : round_up(p->d, 4);
  l.OS = round_up(l.dp4 + 1, 4);
  l.Lp = round_up(p->L, 8) + (lin ? round_up(p->n, 8) : 0);
  l.Lp_ft = l.Lp;
  l.np_ = lin ? 0 : round_up(p->n, 8);
  l.tkb = TP_TKB;
  layout_blocks(l);
  return tp_supported(l);
}
// Pass shape and grid of a launch.  Sample chunks per CTA: 1 (S <= 64: two row tiles per pass), else
// 2 or 4 with one row tile per pass, as many as shared memory allows for this reduction (the
// top-k lists of every chunk live in shared memory); gy CTA columns cover the chunks, gx * gy <= #SM
// (one wave of persistent CTAs).
struct TpPlan { int nch, tiles, gx, gy; };
inline TpPlan tp_plan(const Layout& l, long long N, int sms, int epi, int k) {
  const int chunks = l.Sp / TN;
  int nch = chunks == 1 ? 1 : (chunks == 2 ? 2 : 4);
  while (nch > 1) {
    const TpSmem sl = tp_smem_layout(l, epi, k, nch);
    if (sl.nsb >= (uint32_t)TP_NSB_MIN && sl.total <= TC_SMEM_MAX) break;
    nch /= 2;
  }
  TpPlan p{};
  p.nch = nch;
  p.tiles = nch == 1 ? 2 : 1;
  p.gy = (chunks + nch - 1) / nch;
  const long long nt = (N + p.tiles * TM - 1) / (p.tiles * TM);
  long long gx = sms / p.gy;
  if (gx > nt) gx = nt;
  if (gx < 1) gx = 1;
  p.gx = (int)gx;
  return p;
}
// number of per-CTA partial results of a launch (rows of the [parts][Sp] workspace arrays)
inline int tensor_parts_any(const Layout& l, long long N, int sms, int epi, int k) {
  return l.tc_proj ? tp_plan(l, N, sms, epi, k).gx : tensor_parts(l, N, sms);
}

// floats appended to the tensor pack: one (hi | lo) Omega_aug image per 32-feature block
__host__ __device__ inline size_t tp_pack_floats(const Layout& l) { return (size_t)(l.nkb_r + l.nkb_k) * (tp_obytes(l.d) / 4); }

#ifdef PSBAX_TU_API  // the pack kernel lives in the API translation unit only
// Omega_aug images: UMMA canonical no-swizzle K-major, N = 32 feature rows, K = Kx.
//   float offset inside a split image: (r/8)*(Kx/4)*32 + (k/4)*32 + (r%8)*4 + (k%4)
__global__ void k_tensor_proj_pack(Layout lay, float* __restrict__ pack, size_t off_om) {
  const float* __restrict__ ft = pack + lay.off_ft;
  float* __restrict__ om = pack + off_om;
  const int kx = tp_kx(lay.d), nkb = lay.nkb_r + lay.nkb_k;  // (images of exact-kernel blocks stay zero: no projection)
  const int img = 32 * kx;  // floats per split image
  const size_t total = (size_t)nkb * 2 * img;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int within = (int)(idx % img);
    const int split = (int)((idx / img) & 1);
    const int kb = (int)(idx / (2 * (size_t)img));
    const int rg = within / (kx * 8), rem = within % (kx * 8);
    const int kc = rem / 32, r8 = (rem >> 2) & 7, ke = rem & 3;
    const int r = rg * 8 + r8, k = kc * 4 + ke;
    const int row = kb * TP_TKB + r;  // feature row of the table
    float x = 0.f;
    if (kb < lay.nkb_r && row < lay.Lp_ft) {  // cos / linear rows: (omega | bias at dp4); other rows project to 0
      if (k < lay.d) x = ft[(size_t)row * lay.OS + k];
      else if (k == lay.d && tp_bias_in_x(lay.d)) x = ft[(size_t)row * lay.OS + lay.dp4];
    }
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    om[idx] = split == 0 ? hi : (x - hi);
  }
}

inline int tensor_proj_pack(const Layout& l, float* pack, cudaStream_t st) {
  const size_t off_om = l.off_tc + tensor_pack_floats(l);
  k_tensor_proj_pack<<<148, 256, 0, st>>>(l, pack, off_om);
  return cudaGetLastError() == cudaSuccess ? PSBAX_OK : PSBAX_E_CUDA;
}

#endif  // PSBAX_TU_API

// Waiting roles back off between probes: in this kernel the producers are NOT the bottleneck, and
// twenty lanes re-polling mbarriers at full rate starve the MMA issuer's own barrier traffic in
// the shared-memory sync unit (measured: 4600 cycles per block, MMAs or not).
#ifndef TP_BACKOFF
#define TP_BACKOFF 64
#endif
// The fallback poll is a try_wait with a suspend-time hint: the hardware parks the thread until the
// phase completes or the hint expires.  (__nanosleep between plain try_waits did not hold the pollers
// back: ncu source view of the C5 kernel, round 2 — the four epilogue lanes waiting for d_full with
// __nanosleep(500) still executed a probe every ~50 cycles, 16 % of all issued instructions of a
// kernel whose issue slots are 69 % busy; all poll loops together ~30 %.)
#ifndef TP_WAIT_HINT
#define TP_WAIT_HINT 1  // 0: the nanosleep form
#endif
template <int NS = TP_BACKOFF>
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
#if TP_WAIT_HINT
  constexpr uint32_t HINT = NS >= 200 ? 20000u : 2000u;  // ns
  while (!mbar_try_wait_hint(bar, parity, HINT)) {
    if (++spins > (1u << 22)) asm volatile("trap;");
  }
#else
  while (!mbar_try_wait(bar, parity)) {
    if (NS > 0) __nanosleep(NS);
    if (++spins > (1u << 24)) asm volatile("trap;");
  }
#endif
}
// as mbar_wait_hybrid, with a backed-off fallback poll (the producers of this kernel have slack).
// NS is the sleep between probes: the epilogue waits most of a pass for d_full, and at 64 ns its
// four pollers were 14 % of all issued instructions of an issue-bound kernel (ncu source view).
template <int NS = TP_BACKOFF>
__device__ __forceinline__ void mbar_wait_hybrid_backoff(uint64_t* bar, uint32_t parity, int lane) {
  if (!__all_sync(0xffffffffu, mbar_test(bar, parity))) {
    if (lane == 0) mbar_wait_backoff<NS>(bar, parity);
    __syncwarp();
  }
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// Order of a pass's blocks: pack order (cos / linear blocks, then the exact-kernel blocks), tabulated per
// CTA so that every role walks the same sequence.  TP_INTERLEAVE = 1 alternates runs of `s` cos blocks and
// `t` exact-kernel blocks in the one-tile / several-chunk shape (S > 64), which overlaps the tensor-heavy
// and the FFMA / shared-memory-heavy kind in the pipeline (C3 top-10 1.07 -> 1.00 ms) — and is OFF because
// it costs accuracy: on a Matheron draw ||V_s|| is tens (C3: 50), the partial sums of k(x, X) . V_s are
// large until the last exact-kernel block has cancelled them, and with the cos blocks accumulated in
// between, every one of their (many) MMAs rounds the fp32 accumulators at that magnitude: max value
// error 1.2 x the strict tolerance at C3 against 0.69 x in pack order (scripts/c3_err_probe.py).
// (Two-tile shape: interleaving also lost 4 % at C5 — the contraction is issued in order behind a slow
// exact-kernel block with only two A stages to run ahead in.)
#ifndef TP_PAIRX_MIN_DP4
#define TP_PAIRX_MIN_DP4 8  // packed-FP32 exact-kernel rows over the row pair from this padded d on
#endif
#ifndef TP_INTERLEAVE
#define TP_INTERLEAVE 0
#endif
// The order is tabulated once per CTA in shared memory (uint16 per position: bit 15 = exact-kernel
// block, low bits = image index): the roles then pay one LDS per block (walking the sequence with
// counters cost every warp ~25 instructions per block, 3-5 % of the C5 kernel).
__device__ __forceinline__ uint16_t tp_seq_entry(int p, int nr, int nk, bool il) {
  if (!TP_INTERLEAVE || !il || nr == 0 || nk == 0) return (uint16_t)(p | (p >= nr ? 0x8000 : 0));
  const int s = nr > nk ? nr / nk : 1, t = nk > nr ? nk / nr : 1;
  const int G = min(nr / s, nk / t);  // complete runs of (s cos, t exact-kernel) blocks
  if (p < G * (s + t)) {
    const int g = p / (s + t), o = p - g * (s + t);
    return o < s ? (uint16_t)(g * s + o) : (uint16_t)((nr + g * t + (o - s)) | 0x8000);
  }
  const int q = p - G * (s + t), rl = nr - G * s;  // the surplus: cos blocks first
  return q < rl ? (uint16_t)(G * s + q) : (uint16_t)((nr + G * t + (q - rl)) | 0x8000);
}
struct TpSeq {
  const uint16_t* tab;
  int n, pos;
  __device__ __forceinline__ TpSeq(const uint16_t* tab_, int n_) : tab(tab_), n(n_), pos(0) {}
  // advance to the next block of the sequence: is it an exact-kernel block, and its image index
  __device__ __forceinline__ bool next(int* idx) {
    const uint32_t e = tab[pos];
    pos = (pos + 1 == n) ? 0 : pos + 1;
    *idx = (int)(e & 0x7fffu);
    return (e & 0x8000u) != 0;
  }
};

template <int EPI, bool PAIRX, int NCH>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tensor_proj(const __grid_constant__ KArgs a, size_t off_om) {
  using C = TcCfg<true, true>;  // BF16x3 contraction with 32-feature blocks
  using M = TpCols<NCH>;        // tensor-memory map
  constexpr int TILES = M::TILES;
  constexpr int NSA = M::NSA;
  // Producer groups taking blocks in turn (see the producers): two in the two-tile shape.  In the one-tile /
  // several-chunk shape (S > 64) every warp works on every block: the operand ring there is 3 stages of
  // 42 KB (C3), a block's operands arrive ~2 k cycles after the stage is released, and groups working
  // on different blocks wait for their stage instead of overlapping (tried: 2 and 4 groups with a
  // 4 x 2 / 4 x 4 register-blocked exact-kernel path, C3 1.1 -> 1.4 / 1.7 ms).
  constexpr int GR = NCH == 1 ? 2 : 1;
  static_assert(NSA % GR == 0, "a group owns the A stages of its blocks");
  constexpr int ZB = GR > 2 ? GR : 2;  // Z barrier pairs (see below)
  static_assert(!(PAIRX && TILES != 2), "the packed-FP32 kernel rows pair the two row tiles of a pass");
  extern __shared__ __align__(128) uint8_t tp_smem[];
  uint8_t* const smem = tp_smem;
  const Layout& lay = a.lay;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TpSmem sl = tp_smem_layout(lay, EPI, a.k, NCH);
  uint8_t* sB = smem + sl.b;
  float* sX = reinterpret_cast<float*>(smem + sl.x);
  float* sOut = reinterpret_cast<float*>(smem + sl.out);
  float* sIl = reinterpret_cast<float*>(smem + sl.il);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sl.bars);
  uint64_t* b_full = bars;         // [8]
  uint64_t* b_empty = bars + 8;    // [8]
  uint64_t* a_full = bars + 16;    // [4]
  uint64_t* a_empty = bars + 20;   // [4]
  // Z: two tensor-memory buffers (block it -> buffer it & 1) but ZB = max(2, GR) barrier pairs
  // (block it -> pair it % ZB): a producer group only ever waits on its own pair, so it sees every
  // completion of that barrier in order (with one pair per buffer a group that skips a completion
  // cannot tell "not yet" from "already twice" by the phase parity)
  uint64_t* z_full = bars + 24;    // [4]
  uint64_t* z_empty = bars + 28;   // [4]
  uint64_t* x_full = bars + 32;
  uint64_t* x_empty = bars + 33;
  uint64_t* d_full = bars + 34;
  uint64_t* d_empty = bars + 35;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 36);
  const uint32_t NSB = sl.nsb;

  const int d = lay.d, OS = lay.OS, dp4 = lay.dp4;
  const int kx = tp_kx(d);
  const int nkb_r = lay.nkb_r;              // 32-feature blocks (bf16x3)
  const int nkb = lay.nkb_r + lay.nkb_k;    // + exact-kernel blocks of KF rows (3xTF32), see layout_blocks()
  constexpr int KF = TP_TKB / 2;
  const uint32_t STAGE = sl.stage;
  const uint32_t om_split = 32u * (uint32_t)kx * 4u;  // bytes per split image
  const int chunk0 = blockIdx.y * NCH;                 // first sample chunk of this CTA
  const int s0 = chunk0 * TN;
  const int nch = min(NCH, lay.Sp / TN - chunk0);      // chunks that exist (the last CTA column may have fewer)
  const int npass = (a.ntiles + TILES - 1) / TILES;
  int t0, t1;
  tc_tile_range(npass, gridDim.x, blockIdx.x, &t0, &t1);
  // reduction state of sample chunk c (carved on the fly: an array of these would live in local memory)
  auto epi_of = [&](int c) { return epi_carve(reinterpret_cast<float*>(smem + sl.state + c * sl.state_stride), EPI, a.k); };

  for (int i = tid; i < round_up(dp4, 4); i += TC_THREADS) sIl[i] = __ldg(a.pack + lay.off_invls + i);
  uint16_t* sSeq = reinterpret_cast<uint16_t*>(smem + sl.seq);
  for (int i = tid; i < nkb; i += TC_THREADS) sSeq[i] = tp_seq_entry(i, nkb_r, lay.nkb_k, NCH > 1);
  // x tile of the exact-kernel rows: [i][row] (tile 0, tile 1) pairs for the packed-FP32 path, else
  // [tile][i][row] for the scalar path.  A template parameter, not a runtime branch: at the
  // 80-register cap the mere presence of the second path cost the cos path 10 % (spills).
  constexpr bool pairx = PAIRX;
  if (lay.n > 0) {  // padding dimensions d..dp4-1 are read but never refilled
    if (TILES == 1) {  // [dim pair][row][2]
      for (int i = tid; i < dp4 * TM; i += TC_THREADS) sX[i] = 0.f;
    } else {
      for (int i = tid; i < TILES * (dp4 - d) * TM; i += TC_THREADS) {
        if (pairx) {
          sX[d * TM * TILES + i] = 0.f;
        } else {
          const int t = i / ((dp4 - d) * TM), r = i % ((dp4 - d) * TM);
          sX[(t * dp4 + d) * TM + r] = 0.f;
        }
      }
    }
  }
  const bool bias_in_x = tp_bias_in_x(d);
  float* sBias = reinterpret_cast<float*>(smem + sl.bias);
  if (!bias_in_x)
    for (int i = tid; i < lay.Kp; i += TC_THREADS)
      sBias[i] = i < lay.Lp_ft ? __ldg(a.pack + lay.off_ft + (size_t)i * OS + dp4) : 0.f;
  if (tid == 0) {
    // the producers read the streamed table block, so they release the stage too
    // a block is produced by one group of TC_NPROD / GR warps (see the producers)
    for (int i = 0; i < TP_NSB_MAX; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1 + TC_NPROD / GR); }
    for (int i = 0; i < NSA; ++i) { mbar_init(&a_full[i], TC_NPROD / GR); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < ZB; ++i) { mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], TC_NPROD / GR); }
    mbar_init(x_full, TC_NPROD); mbar_init(x_empty, 1);
    mbar_init(d_full, 1);        mbar_init(d_empty, TC_NEPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_WARP_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_NEPI) {
    for (int c = 0; c < NCH; ++c) epi_init(epi_of(c), EPI, a.k, tid - TC_EPI_WARP0 * 32, TC_NEPI * 32);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t total_blocks = (uint32_t)(t1 - t0) * (uint32_t)nkb;

  if (warp == TC_WARP_LOAD) {
    // ================= loader: the chunks' [W;V] images + Omega_aug images of each block =================
    const uint8_t* gW = reinterpret_cast<const uint8_t*>(a.pack + lay.off_tc) + (size_t)chunk0 * nkb * TP_WBYTES;
    const uint8_t* gO = reinterpret_cast<const uint8_t*>(a.pack + off_om);
    const uint8_t* gFT = reinterpret_cast<const uint8_t*>(a.pack + lay.off_ft);
    const uint32_t obytes = tp_obytes(d);
    const uint32_t wbytes = (uint32_t)nch * TP_WBYTES;
    uint32_t bs = 0, ph = 0;  // ring stage and its phase
    TpSeq seq(sSeq, nkb);
    for (uint32_t it = 0; it < total_blocks; ++it) {
      int kb;  // image index of this block
      const bool kblk = seq.next(&kb);  // exact-kernel block: [W;V] images (tf32) + its KF table rows, no projection
      TC_TRACE(3, it, 0);
      mbar_wait_hybrid_backoff<200>(&b_empty[bs], ph ^ 1, lane);
      TC_TRACE(3, it, 1);
      if (elect_one()) {
        const uint32_t fb = (uint32_t)KF * OS * 4;
        mbar_arrive_expect_tx(&b_full[bs], wbytes + (kblk ? fb : obytes));
        for (int c = 0; c < nch; ++c)  // image (chunk0 + c, kb)
          bulk_g2s(sB + bs * STAGE + c * TP_WBYTES, gW + ((size_t)c * nkb + kb) * TP_WBYTES, TP_WBYTES, &b_full[bs]);
        if (kblk)
          bulk_g2s(sB + bs * STAGE + sl.ftk, gFT + ((size_t)lay.kk0 + (size_t)(kb - nkb_r) * KF) * OS * 4, fb, &b_full[bs]);
        else
          bulk_g2s(sB + bs * STAGE + NCH * TP_WBYTES, gO + (size_t)kb * obytes, obytes, &b_full[bs]);
      }
      __syncwarp();
      if (++bs == NSB) { bs = 0; ph ^= 1; }
    }
  } else if (warp == TP_WARP_PROJ) {
    // ================= projection issuer: Z[jt & 1] = X_aug * Omega_aug(block jt)^T =================
    // The two MMA streams have their own warps: a single warp walking both wait -> issue ->
    // commit chains was the critical path (~1800 cycles per block against ~1100 for the
    // producers).  All lanes poll: a lane-0 poll leaves the warp diverged, and every
    // reconvergence (WARPSYNC slow path) costs a latency-critical warp a few hundred cycles
    // (measured 4600 -> 1800 cycles per block).  Counters instead of divisions, descriptors by
    // addition from per-kernel bases.
    constexpr uint32_t IDESC_P = umma_idesc_tf32(32);
    const uint32_t sbo_o = (uint32_t)(kx / 4) * 128u;
    // the 14-bit start-address field counts 16-byte units; stage offsets never carry out of it
    const uint64_t dO0 = umma_desc_kmajor(smem_u32(sB) + NCH * TP_WBYTES, 128, sbo_o);
    const uint32_t stage16 = STAGE >> 4, osplit16 = om_split >> 4;
    const int nks = kx / 8;
    const uint32_t nkb1 = (uint32_t)nkb - 1;
    uint32_t jkb = 0, pl = 0, bs = 0, bph = 0;
    TpSeq seq(sSeq, nkb);
    for (uint32_t jt = 0; jt < total_blocks; ++jt) {
      const uint32_t zb = jt & 1;
      int bidx;
      const bool kblk = seq.next(&bidx);
      if (jkb == 0) mbar_wait(x_full, pl & 1);  // this pass's X operand is in tensor memory
      mbar_wait(&b_full[bs], bph);
      if (jt >= 2) mbar_wait(&z_empty[(jt - 2) % ZB], ((jt - 2) / ZB) & 1);  // block jt - 2 has left this Z buffer
      __syncwarp();
      tc_fence_after();
      TC_TRACE(0, jt, 4);
      if (elect_one()) {
        const uint64_t oh0 = dO0 + (uint64_t)(bs * stage16), ol0 = oh0 + osplit16;
        // exact-kernel blocks have no projection: only the barrier protocol runs.  Consecutive MMAs alternate
        // between the row tiles' Z accumulators (a run of MMAs into one accumulator is serialised by the
        // tensor pipe: sp_tensor.cuh, issue order); per accumulator the order of additions is unchanged.
        if (!kblk) {
#pragma unroll 1
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t oh = oh0 + (uint64_t)(ks * 16), ol = ol0 + (uint64_t)(ks * 16);
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int t = 0; t < TILES; ++t) {
                const uint32_t dz = tmem + M::Z + (zb * TILES + t) * 32;
                const uint32_t xh = tmem + M::X + (t * 2) * 32 + ks * 8, xl = xh + 32;
                tc_mma_ts(dz, pr == 2 ? xl : xh, pr == 1 ? ol : oh, IDESC_P, (ks | pr) ? 1u : 0u);
              }
          }
        }
        tc_commit(&z_full[jt % ZB]);
        if (jkb == nkb1) tc_commit(x_empty);  // last reader of this pass's X
      }
      __syncwarp();
      TC_TRACE(0, jt, 5);
      if (jkb == nkb1) { jkb = 0; ++pl; } else ++jkb;
      if (++bs == NSB) { bs = 0; bph ^= 1; }
    }
  } else if (warp == TC_WARP_MMA) {
    // ================= contraction issuer: D[chunk][tile] += A(block it) * [W;V](chunk, block it)^T =================
    constexpr uint32_t IDESC_C = umma_idesc_bf16(TN);
    const uint64_t dW0 = umma_desc_kmajor(smem_u32(sB), 128, C::B_SBO);
    const uint32_t stage16 = STAGE >> 4;
    const uint32_t nkb1 = (uint32_t)nkb - 1;
    uint32_t kb = 0, pl = 0, bs = 0, bph = 0;  // kb: position inside the pass
    TpSeq seq(sSeq, nkb);
    for (uint32_t it = 0; it < total_blocks; ++it) {
      const bool last = kb == nkb1;
      int bidx;
      const bool kblk = seq.next(&bidx);
      const uint32_t as = it % NSA;
      TC_TRACE(0, it, 0);
      if (kb == 0) mbar_wait(d_empty, (pl & 1) ^ 1);
      mbar_wait(&b_full[bs], bph);  // long complete (the projection of this block read the stage); acquire only
      mbar_wait(&a_full[as], (it / NSA) & 1);
      __syncwarp();
      tc_fence_after();
      TC_TRACE(0, it, 2);
      if (elect_one()) {
        const uint64_t b10 = dW0 + (uint64_t)(bs * stage16);
        if (kblk) {
          // exact-kernel block: KF features as TF32 hi | lo (A columns [0, 16) | [16, 32)), K = 8 per MMA
          constexpr uint32_t IDESC_K = umma_idesc_tf32(TN);
          constexpr uint32_t SBO_K = KF * 4 * 8, SPLIT2_K = 8 * SBO_K;
          const uint64_t k10 = umma_desc_kmajor(smem_u32(sB), 128, SBO_K) + (uint64_t)(bs * stage16);
          // k-step major, then the three split products, the (chunk, tile) accumulators innermost: consecutive
          // MMAs go to different accumulators; per accumulator the order of additions is unchanged
#pragma unroll
          for (int ks = 0; ks < KF / 8; ++ks)
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int t = 0; t < TILES; ++t) {
                  if (c >= nch) continue;  // (the last CTA column may have fewer chunks)
                  const uint64_t b1 = k10 + (uint64_t)(c * (TP_WBYTES >> 4) + ks * 16);
                  const uint64_t b2 = b1 + (uint64_t)(SPLIT2_K >> 4);
                  const uint32_t a1 = tmem + M::A + (as * TILES + t) * 32 + ks * 8, a2 = a1 + 16;
                  const uint32_t dd = tmem + M::D + (c * TILES + t) * TN;
                  tc_mma_ts(dd, pr == 2 ? a2 : a1, pr == 1 ? b2 : b1, IDESC_K, (kb | ks | pr) ? 1u : 0u);
                }
        } else {
#pragma unroll
          for (int ks = 0; ks < C::KSTEPS; ++ks)
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int t = 0; t < TILES; ++t) {
                  if (c >= nch) continue;
                  const uint64_t b1 = b10 + (uint64_t)(c * (TP_WBYTES >> 4) + ks * (C::B_KSTEP >> 4));
                  const uint64_t b2 = b1 + (uint64_t)(C::B_SPLIT2 >> 4);
                  const uint32_t a1 = tmem + M::A + (as * TILES + t) * 32 + ks * 8, a2 = a1 + 16;
                  const uint32_t dd = tmem + M::D + (c * TILES + t) * TN;
                  tc_mma_ts_bf16(dd, pr == 2 ? a2 : a1, pr == 1 ? b2 : b1, IDESC_C, (kb | ks | pr) ? 1u : 0u);
                }
        }
        tc_commit(&a_empty[as]);
        tc_commit(&b_empty[bs]);
        if (last) tc_commit(d_full);
      }
      __syncwarp();
      TC_TRACE(0, it, 3);
      if (last) { kb = 0; ++pl; } else ++kb;
      if (++bs == NSB) { bs = 0; bph ^= 1; }
    }
  } else if (warp >= TC_PROD_WARP0 && warp < TC_PROD_WARP0 + TC_NPROD) {
    // ================= producers =================
    // Two groups of eight warps (two per lane quarter) take ALTERNATE blocks: group g owns the blocks
    // with (it & 1) == g, i.e. Z buffer g and the A stages of parity g, and a warp evaluates 16
    // features (8 of an exact-kernel block) of its block.  With all sixteen warps on every block
    // (8 features each) a block's fixed latencies — z_full wait, tcgen05.ld, a_empty wait, tcgen05.st,
    // wait::st, arrive: ~850 cycles — sat on every warp's path for 512 cycles' worth of MUFU per
    // scheduler, and the warps of a scheduler went through them in step (clock trace, d = 10:
    // 1450 cycles per block).  Now one group waits while the other evaluates.
    const int q = warp & 3, h = (warp - TC_PROD_WARP0) >> 2;  // h: which X part this warp writes
    // (GR = 1 in the one-tile / several-chunk shape: there the blocks are tensor-bound and a warp's
    // 8 features x 4 A stages already cover the latencies; measured 1.09 -> 1.22 ms at C3 with two groups)
    const int g = h % GR, h2 = h / GR;                       // group; which part of the block's features
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    constexpr int PT = TC_NPROD * 32;
    const int ptid = tid - TC_PROD_WARP0 * 32;
    const bool tracer = q == 0 && h2 == 0;  // warp 0 (and, with two groups, warp 4 for its blocks)
    (void)tracer;
    uint32_t it = 0, bsc = 0, bphc = 0;
    TpSeq seq(sSeq, nkb);
    for (int pass = t0; pass < t1; ++pass) {
      const uint32_t pl = (uint32_t)(pass - t0);
      // ---- X operand of this pass: warp h writes tile (h >> 1), split (h & 1) of its rows ----
      mbar_wait_hybrid_backoff(x_empty, (pl & 1) ^ 1, lane);
      tc_fence_after();
      if (h < 2 * TILES) {
        const int t = h >> 1, split = h & 1;
        const long long row = ((long long)pass * TILES + t) * TM + q * 32 + lane;
        const float* xr = a.X + row * d;
        const bool ok = row < a.N;
        for (int c = 0; c < kx / 8; ++c) {
          uint32_t v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = c * 8 + j;
            const float xv = (k < d) ? (ok ? __ldg(xr + k) : 0.f) : (k == d ? 1.0f : 0.f);
            const uint32_t hb = (__float_as_uint(xv) + 0x1000u) & 0xffffe000u;  // round to TF32: |lo| <= 2^-11 |x|
            v[j] = split == 0 ? hb : __float_as_uint(xv - __uint_as_float(hb));
          }
          tc_st8(lane_addr + M::X + (t * 2 + split) * 32 + c * 8, v);
        }
        tc_wait_st();
      }
      if (lay.n > 0) {  // plain x tile for the exact-kernel features
        asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");
        const long long base = (long long)pass * TILES * TM * d;
        const long long lim = a.N * d;
        for (int idx = ptid; idx < TILES * TM * d; idx += PT) {
          const int rr = idx / d, i = idx - rr * d;
          const float xv = (base + idx < lim) ? __ldg(a.X + base + idx) : 0.f;
          if (pairx) sX[(i * TM + (rr & 127)) * TILES + (rr >> 7)] = xv;  // one LDS.64 feeds a packed FFMA2
          else if (TILES == 1) sX[((i >> 1) * TM + rr) * 2 + (i & 1)] = xv * sIl[i];  // scaled, [dim pair][row][2]
          else sX[((rr >> 7) * dp4 + i) * TM + (rr & 127)] = xv;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_full);

      for (int pos = 0; pos < nkb; ++pos, ++it) {
        const uint32_t bs = bsc, bph = bphc;  // ring stage of block it (every warp counts every block)
        if (++bsc == NSB) { bsc = 0; bphc ^= 1; }
        int kb;  // image index of this block
        const bool kblk = seq.next(&kb);  // exact-kernel block (KF rows, 3xTF32): no projection arguments to read
        if (GR > 1 && (it & (uint32_t)(GR - 1)) != (uint32_t)g) continue;
        const uint32_t as = it % NSA, aph = (it / NSA) & 1, zb = it & 1, zi = it % ZB, zph = (it / ZB) & 1;
        const int kbase = kb * TP_TKB + h2 * (8 * GR);
        if (tracer) TC_TRACE(1, it, 0);
        mbar_wait_hybrid_backoff(&z_full[zi], zph, lane);
        tc_fence_after();
        if (tracer) TC_TRACE(1, it, 1);
        if (!kblk) {
          // ---- arguments of this warp's 16 features x TILES tiles from the projection ----
          uint32_t z[GR][TILES][8];
#pragma unroll
          for (int sg = 0; sg < GR; ++sg)
#pragma unroll
            for (int t = 0; t < TILES; ++t)
              tc_ld8(lane_addr + M::Z + (zb * TILES + t) * 32 + h2 * (8 * GR) + sg * 8, z[sg][t]);
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&z_empty[zi]);
            mbar_arrive(&b_empty[bs]);  // (cos / linear blocks read nothing from the operand stage)
          }
          if (tracer) TC_TRACE(1, it, 2);
#pragma unroll
          for (int sg = 0; sg < GR; ++sg) {
            const int kb8 = kbase + sg * 8;
            float v[TILES][8];
            if (kb8 < lay.Lp_ft) {
              const bool lin = kb8 >= lay.Lcos;
              float arg[TILES][8];
#pragma unroll
              for (int t = 0; t < TILES; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) arg[t][j] = __uint_as_float(z[sg][t][j]);
              if (!bias_in_x) {  // d % 8 == 0: the bias is not part of the projection
                const float4 b0 = *reinterpret_cast<const float4*>(sBias + kb8);
                const float4 b1 = *reinterpret_cast<const float4*>(sBias + kb8 + 4);
                const float bj[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int t = 0; t < TILES; ++t)
#pragma unroll
                  for (int j = 0; j < 8; ++j) arg[t][j] += bj[j];
              }
#pragma unroll
              for (int t = 0; t < TILES; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[t][j] = lin ? arg[t][j] : __cosf(arg[t][j]);
            } else {  // padding rows between the RFF section and the first exact-kernel block (zero weights)
#pragma unroll
              for (int t = 0; t < TILES; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[t][j] = 0.f;
            }
            // ---- BF16x3 split: 8 features as bf16 pairs (split 1 | split 2) ----
            uint32_t s1[TILES][4], s2[TILES][4];
#pragma unroll
            for (int t = 0; t < TILES; ++t)
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t p1 = cvt_bf16x2(v[t][2 * c + 1], v[t][2 * c]);
                float r0, r1;
#if PSBAX_SPLIT_FHFMA
                bf16_residual2(p1, v[t][2 * c], v[t][2 * c + 1], r0, r1);
#else
                const uint64_t rr = ffma2(pack2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u)),
                                          pack2(-1.0f, -1.0f), pack2(v[t][2 * c], v[t][2 * c + 1]));
                unpack2(rr, r0, r1);
#endif
                s1[t][c] = p1;
                s2[t][c] = cvt_bf16x2(r1, r0);
              }
            if (sg == 0) {
              if (tracer) TC_TRACE(1, it, 3);
              mbar_wait_hybrid_backoff(&a_empty[as], aph ^ 1, lane);
              tc_fence_after();
              if (tracer) TC_TRACE(1, it, 4);
            }
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
              const uint32_t col = M::A + (as * TILES + t) * 32 + h2 * (4 * GR) + sg * 4;
              tc_st4(lane_addr + col, s1[t]);
              tc_st4(lane_addr + col + 16, s2[t]);
            }
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&z_empty[zi]);
          if (tracer) TC_TRACE(1, it, 2);
          // ---- exact-kernel block: KF = 16 features, KW = 4 GR per warp (FFMA path: r^2 as a sum of
          // squared differences), TF32 hi | lo operands.  (All sixteen warps on these blocks, with the
          // groups alternating on the others only, was slower: nobody works ahead of such a block.) ----
          // the block's table rows sit behind the [W;V] images of the same stage
          constexpr int KW = 4 * GR;
          mbar_wait_hybrid_backoff(&b_full[bs], bph, lane);
          const float* ft = reinterpret_cast<const float*>(sB + bs * STAGE + sl.ftk) + (size_t)(h2 * KW) * OS;
          float acc[TILES][KW];
          if constexpr (PAIRX) {
            // packed FP32 over the thread's (tile 0, tile 1) rows; inv_ls and the table entry enter as
            // scalar-broadcast operands:  df = x * inv_ls + (-x_t * inv_ls),  acc += df * df
            const uint64_t* xr = reinterpret_cast<const uint64_t*>(sX) + q * 32 + lane;
            uint64_t acc2[KW];
#pragma unroll
            for (int j = 0; j < KW; ++j) acc2[j] = 0ull;
            for (int i0 = 0; i0 < dp4; i0 += 4) {
              uint64_t xa[4];
              float il4[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                xa[c] = xr[(i0 + c) * TM];
                il4[c] = sIl[i0 + c];
              }
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
            for (int j = 0; j < KW; ++j) unpack2(acc2[j], acc[0][j], acc[TILES - 1][j]);
          } else if constexpr (TILES == 1) {
            // one row per thread: packed FP32 over PAIRS OF DIMENSIONS.  The scaled row sits in shared
            // memory as (x_i, x_i+1) / ls pairs, a table row is read as 16-byte pieces = two pairs:
            //   df = (x_i, x_i+1) / ls + (-x_t)[i, i+1] / ls,  (acc_even, acc_odd) += df * df,  r^2 = acc_even + acc_odd
            const uint64_t* xs2 = reinterpret_cast<const uint64_t*>(sX) + q * 32 + lane;
            uint64_t acc2[KW];
#pragma unroll
            for (int j = 0; j < KW; ++j) acc2[j] = 0ull;
            const uint64_t one2 = pack2(1.0f, 1.0f);
            for (int i0 = 0; i0 < dp4; i0 += 4) {
              const uint64_t xa = xs2[(i0 >> 1) * TM], xb = xs2[((i0 >> 1) + 1) * TM];
#pragma unroll
              for (int j = 0; j < KW; ++j) {
                const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(ft + j * OS + i0);
                const uint64_t dfa = ffma2(xa, one2, w.x), dfb = ffma2(xb, one2, w.y);
                acc2[j] = ffma2(dfa, dfa, acc2[j]);
                acc2[j] = ffma2(dfb, dfb, acc2[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < KW; ++j) {
              float ev, od;
              unpack2(acc2[j], ev, od);
              acc[0][j] = ev + od;
            }
          } else {
            const float* xs = sX + q * 32 + lane;
#pragma unroll
            for (int t = 0; t < TILES; ++t)
#pragma unroll
              for (int j = 0; j < KW; ++j) acc[t][j] = 0.f;
            for (int i0 = 0; i0 < dp4; i0 += 4) {
              float xa[TILES][4];
#pragma unroll
              for (int t = 0; t < TILES; ++t)
#pragma unroll
                for (int c = 0; c < 4; ++c) xa[t][c] = xs[(t * dp4 + i0 + c) * TM] * sIl[i0 + c];
#pragma unroll
              for (int j = 0; j < KW; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(ft + j * OS + i0);
                const float wv4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int t = 0; t < TILES; ++t)
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const float df = xa[t][c] + wv4[c];
                    acc[t][j] = fmaf(df, df, acc[t][j]);
                  }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&b_empty[bs]);
          if (tracer) TC_TRACE(1, it, 3);
          mbar_wait_hybrid_backoff(&a_empty[as], aph ^ 1, lane);
          tc_fence_after();
          if (tracer) TC_TRACE(1, it, 4);
#pragma unroll
          for (int t = 0; t < TILES; ++t) {
            float kv[KW];
            kernel_profile_n<KW>(acc[t], kv, lay.kernel);
            uint32_t khi[KW], klo[KW];
#pragma unroll
            for (int j = 0; j < KW; ++j) {
              const uint32_t hb = (__float_as_uint(kv[j]) + 0x1000u) & 0xffffe000u;  // round to TF32: |lo| <= 2^-11 |x|, unbiased
              khi[j] = hb;
              klo[j] = __float_as_uint(kv[j] - __uint_as_float(hb));
            }
            const uint32_t col = M::A + (as * TILES + t) * 32 + h2 * KW;
            if constexpr (KW == 16) {
              tc_st16(lane_addr + col, khi);
              tc_st16(lane_addr + col + 16, klo);
            } else if constexpr (KW == 8) {
              tc_st8(lane_addr + col, khi);
              tc_st8(lane_addr + col + 16, klo);
            } else {
              tc_st4(lane_addr + col, khi);
              tc_st4(lane_addr + col + 16, klo);
            }
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (tracer) TC_TRACE(1, it, 5);
      }
    }
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + TC_NEPI) {
    // ================= epilogue (D single-buffered) =================
    const int q = warp & 3;
    const int etid = tid - TC_EPI_WARP0 * 32;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int r = q * 32 + lane;
    if constexpr (NCH == 1) {
      // drain both tiles, release the accumulators, then reduce (level set: register-resident state)
      LsRegState<TC_NEPI> ls;
      ls.init();
      for (int pass = t0; pass < t1; ++pass) {
        const uint32_t pl = (uint32_t)(pass - t0);
        mbar_wait_hybrid_backoff<500>(d_full, pl & 1, lane);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < TILES; ++t) {
#pragma unroll 1
          for (int c = 0; c < TN / 16; ++c) {
            uint32_t u[16];
            tc_ld16(lane_addr + M::D + t * TN + c * 16, u);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) sOut[(t * TN + c * 16 + j) * OUT_LD + r] = __uint_as_float(u[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
#pragma unroll 1
        for (int t = 0; t < TILES; ++t) {
          const int tile = pass * TILES + t;
          if (tile < a.ntiles) {
            const float* so = sOut + (size_t)t * TN * OUT_LD;
            if constexpr (EPI == EPI_LEVELSET) epilogue_levelset_reg<TC_NEPI>(a, so, ls, tile * TM, s0, etid);
            else epilogue_tile<EPI, TC_NEPI>(a, so, epi_of(0), tile * TM, s0, etid);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
      }
      if constexpr (EPI == EPI_LEVELSET) epilogue_levelset_reg_flush<TC_NEPI>(a, ls, s0, etid);
      else epilogue_flush<EPI>(a, epi_of(0), s0, etid, TC_NEPI * 32);
    } else {
      // one tile, NCH sample chunks: drain and reduce chunk by chunk through one staging tile; the
      // accumulators are released as soon as the last chunk has left tensor memory (the reduction
      // state of every chunk lives in shared memory)
      for (int pass = t0; pass < t1; ++pass) {
        const uint32_t pl = (uint32_t)(pass - t0);
        if (etid == 0) TC_TRACE(2, pl, 0);
        mbar_wait_hybrid_backoff<500>(d_full, pl & 1, lane);
        tc_fence_after();
        if (etid == 0) TC_TRACE(2, pl, 1);
#pragma unroll 1
        for (int ch = 0; ch < nch; ++ch) {
#pragma unroll 1
          for (int c = 0; c < TN / 16; ++c) {
            uint32_t u[16];
            tc_ld16(lane_addr + M::D + ch * TN + c * 16, u);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) sOut[(c * 16 + j) * OUT_LD + r] = __uint_as_float(u[j]);
          }
          if (ch == nch - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty);
            if (etid == 0) TC_TRACE(2, pl, 2);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
          epilogue_tile<EPI, TC_NEPI>(a, sOut, epi_of(ch), pass * TM, s0 + ch * TN, etid);
          asm volatile("bar.sync 1, %0;" ::"n"(TC_NEPI * 32) : "memory");
        }
        if (etid == 0) TC_TRACE(2, pl, 3);
      }
      for (int ch = 0; ch < nch; ++ch) epilogue_flush<EPI>(a, epi_of(ch), s0 + ch * TN, etid, TC_NEPI * 32);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_ALLOC) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
}

template <int EPI, bool PAIRX, int NCH>
int tensor_proj_launch_epi(const KArgs& a, const TpPlan& plan, cudaStream_t st, char* err, size_t errlen) {
  const Layout& l = a.lay;
  const TpSmem sl = tp_smem_layout(l, EPI, a.k, NCH);
  if (sl.total > TC_SMEM_MAX || sl.nsb < (uint32_t)TP_NSB_MIN) {
    snprintf(err, errlen, "tensor projection engine: shared memory %u B exceeds 227 KB", sl.total);
    return PSBAX_E_UNSUPPORTED;
  }
  const size_t off_om = l.off_tc + tensor_pack_floats(l);
  cudaError_t ce = cudaFuncSetAttribute(k_tensor_proj<EPI, PAIRX, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total);
  if (ce == cudaSuccess) {
    k_tensor_proj<EPI, PAIRX, NCH><<<dim3(plan.gx, plan.gy), TC_THREADS, sl.total, st>>>(a, off_om);
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "tensor projection engine launch failed: %s", cudaGetErrorString(ce));
    return PSBAX_E_CUDA;
  }
  return PSBAX_OK;
}

int tensor_proj_launch(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen);  // tensor_proj.cu
#ifdef PSBAX_TU_PROJ
template <int EPI>
int tensor_proj_launch_shape(const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  const TpPlan plan = tp_plan(a.lay, a.N, sms, EPI, a.k);
  if (plan.nch == 4) return tensor_proj_launch_epi<EPI, false, 4>(a, plan, st, err, errlen);
  if (plan.nch == 2) return tensor_proj_launch_epi<EPI, false, 2>(a, plan, st, err, errlen);
  // Two row tiles per pass.  Packed-FP32 exact-kernel rows over the row pair: measured per
  // instantiation (same box, d = 5 / 10 / 32).  The level-set kernel (register-resident epilogue
  // state, tightest register budget) only gains from d = 13 on (d = 10: 7.78 -> 8.23 ms, d = 32:
  // 3.01 -> 2.87); top-k and values gain from d = 5 (C5 top-32 8.65 -> 8.48 ms).
  const bool pairx = a.lay.n > 0 && a.lay.dp4 >= TP_PAIRX_MIN_DP4;
  return pairx ? tensor_proj_launch_epi<EPI, true, 1>(a, plan, st, err, errlen)
               : tensor_proj_launch_epi<EPI, false, 1>(a, plan, st, err, errlen);
}
inline int tensor_proj_launch_impl(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  if (epi == EPI_VALUES) return tensor_proj_launch_shape<EPI_VALUES>(a, sms, st, err, errlen);
  if (epi == EPI_LEVELSET) return tensor_proj_launch_shape<EPI_LEVELSET>(a, sms, st, err, errlen);
  return tensor_proj_launch_shape<EPI_TOPK>(a, sms, st, err, errlen);
}

#endif

inline int tensor_launch_any(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  return a.lay.tc_proj ? tensor_proj_launch(epi, a, sms, st, err, errlen) : tensor_launch(epi, a, sms, st, err, errlen);
}

}  // namespace psbax
