// FFMA + MUFU kernels: fused sample-path evaluation + BAX reduction.
//
//   k_shared_simt : shared basis (G == 1).  Per 128-row tile the RFF / kernel
//                   features are generated once into shared memory and contracted
//                   against a [TK x 64] block of [W;V] with an 8x4 register tile.
//   k_persample   : one basis per sample (G == S, the reference's get_gp_samples
//                   form) and small-S shapes.  lane = candidate row (x in
//                   registers), basis rows broadcast to the warp; MUFU bound.
//
// Both stage the finished [64 samples][128 rows] value tile in shared memory and
// hand it to the same epilogue: values dump, level-set (ballot mask + count +
// arg max) or per-sample top-k (threshold filter + warp insertion).  The N x S
// value matrix only reaches HBM for EPI_VALUES.
#pragma once
#include "sp_common.cuh"

namespace psbax {

// ---------------------------------------------------------------------------------
// per-CTA reduction state (lives in shared memory, outside the aliased tile region)
// ---------------------------------------------------------------------------------
struct EpiSmem {
  unsigned int* cnt;  // [TN]
  float* mx;          // [TN]
  int* arg;           // [TN]
  float* tv;          // [TN][k]
  int* ti;            // [TN][k]
};

__host__ __device__ inline size_t epi_smem_floats(int epi, int k) {
  if (epi == EPI_LEVELSET) return 3 * TN;
  if (epi == EPI_TOPK) return 2 * (size_t)TN * k;
  return 0;
}

__device__ __forceinline__ EpiSmem epi_carve(float* base, int epi, int k) {
  EpiSmem e{};
  if (epi == EPI_LEVELSET) {
    e.cnt = reinterpret_cast<unsigned int*>(base);
    e.mx = base + TN;
    e.arg = reinterpret_cast<int*>(base + 2 * TN);
  } else if (epi == EPI_TOPK) {
    e.tv = base;
    e.ti = reinterpret_cast<int*>(base + TN * k);
  }
  return e;
}

__device__ __forceinline__ void epi_init(const EpiSmem& e, int epi, int k, int tid,
                                         int nthreads = NTHREADS) {
  if (epi == EPI_LEVELSET) {
    for (int i = tid; i < TN; i += nthreads) {
      e.cnt[i] = 0u;
      e.mx[i] = -CUDART_INF_F;
      e.arg[i] = 0x7fffffff;
    }
  } else if (epi == EPI_TOPK) {
    for (int i = tid; i < TN * k; i += nthreads) {
      e.tv[i] = -CUDART_INF_F;
      e.ti[i] = 0x7fffffff;
    }
  }
}

// Merge the candidates of one 32-row group into the sorted (best first) k-list of one sample.
// Called by a full warp; lane j of the group holds candidate (y, row) when bit j of `m` is set.
//
// Rank-based: the list entries (lane-distributed: lane j owns entries j and j + 32) and the
// candidates are merged in ONE pass over the candidates — each iteration broadcasts one candidate
// and every lane updates the new rank of its own list entries and of its own candidate; the
// iterations are independent (no store -> load round trip per candidate), and every element then
// writes itself to its new position (measured on the C3 shape: a cold list cost ~340 cycles per
// serial insertion, ~150 per candidate this way; what removes the cold start itself is the seed
// threshold, see psbax_eval_topk).
// The order is the total order of better(): larger value first, ties to the smaller index.
__device__ __forceinline__ void topk_merge_group(float* tv, int* ti, int k, float y, int row,
                                                 unsigned int m, int lane) {
  const int q0 = lane, q1 = lane + 32;
  const bool two = k > 32;  // warp-uniform
  float e0v = -CUDART_INF_F, e1v = -CUDART_INF_F;
  int e0i = 0x7fffffff, e1i = 0x7fffffff;
  if (q0 < k) { e0v = tv[q0]; e0i = ti[q0]; }
  if (two && q1 < k) { e1v = tv[q1]; e1i = ti[q1]; }
  const bool mine = (m >> lane) & 1u;
  int r0 = q0, r1 = q1;  // new ranks of this lane's list entries
  int rc = 0;            // new rank of this lane's candidate
  unsigned int mm = m;
  while (mm) {
    const int src = __ffs(mm) - 1;
    mm &= mm - 1;
    const float cv = __shfl_sync(0xffffffffu, y, src);
    const int ci = __shfl_sync(0xffffffffu, row, src);
    const bool b0 = better(e0v, e0i, cv, ci);  // (unused slots hold (-inf, int max): never better)
    int nb = __popc(__ballot_sync(0xffffffffu, b0));
    if (!b0) ++r0;  // the candidate goes in front of this entry
    if (two) {
      const bool b1 = better(e1v, e1i, cv, ci);
      nb += __popc(__ballot_sync(0xffffffffu, b1));
      if (!b1) ++r1;
    }
    if (mine && better(cv, ci, y, row)) ++rc;  // candidates in front of mine
    if (lane == src) rc += nb;                 // + list entries in front of mine
  }
  __syncwarp();  // every lane holds its old entries in registers: positions may now be overwritten
  if (q0 < k && r0 < k && r0 != q0) { tv[r0] = e0v; ti[r0] = e0i; }
  if (two && q1 < k && r1 < k && r1 != q1) { tv[r1] = e1v; ti[r1] = e1i; }
  if (mine && rc < k) { tv[rc] = y; ti[rc] = row; }
  __syncwarp();
}

// ---------------------------------------------------------------------------------
// Epilogue over one staged tile.  sOut[sl * OUT_LD + r] holds the raw contraction
// f (before the affine map) of local sample sl, tile row r.
// ---------------------------------------------------------------------------------
// NW warps take part (tid in [0, 32 * NW)).
template <int EPI, int NW = NWARPS>
__device__ __forceinline__ void epilogue_tile(const KArgs& a, const float* sOut,
                                              const EpiSmem& e, int row0, int s0, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const long long N = a.N;
  if constexpr (EPI == EPI_VALUES) {
    if (a.rowsq != nullptr) {  // sum of squares over this chunk's samples, thread <-> row, ascending s
      const int ns = min(TN, a.lay.S - s0);
      const int grp = a.rowsq_group > 0 ? a.rowsq_group : TN;
      for (int r = tid; r < TM; r += NW * 32) {
        const long long row = (long long)row0 + r;
        if (row >= N) break;
        for (int g0 = 0; g0 < ns; g0 += grp) {
          float acc = 0.f;
          const int g1 = min(ns, g0 + grp);
          for (int sl = g0; sl < g1; ++sl) {
            const float y = fmaf(sOut[sl * OUT_LD + r], a.ystd, a.c0);
            acc = fmaf(y, y, acc);
          }
          a.rowsq[(long long)((s0 + g0) / grp) * N + row] = acc;
        }
      }
      return;
    }
    // lanes <-> samples so that global stores are contiguous along S
    for (int r = warp; r < TM; r += NW) {
      const long long row = (long long)row0 + r;
      if (row >= N) break;
#pragma unroll
      for (int h = 0; h < TN / 32; ++h) {
        const int sl = lane + 32 * h;
        const int s = s0 + sl;
        if (s < a.lay.S) a.out[row * a.ld_out + s] = fmaf(sOut[sl * OUT_LD + r], a.ystd, a.c0);
      }
    }
  } else {
  // warp w owns local samples w*(TN/NW) .. ; lanes <-> rows
#pragma unroll 1
  for (int q = 0; q < TN / NW; ++q) {
    const int sl = warp * (TN / NW) + q;
    const int s = s0 + sl;
    if (s >= a.lay.S) break;
    if constexpr (EPI == EPI_LEVELSET) {
      float bv = -CUDART_INF_F;
      int bi = 0x7fffffff;
      unsigned int c = 0;
#pragma unroll
      for (int g = 0; g < TM / 32; ++g) {
        const int r = g * 32 + lane;
        const long long row = (long long)row0 + r;
        const bool valid = row < N;
        const float y = fmaf(sOut[sl * OUT_LD + r], a.ystd, a.c0);
        const unsigned int word = __ballot_sync(0xffffffffu, valid && (y > a.tau));
        c += __popc(word);
        if (a.mask != nullptr && lane == 0 && (long long)row0 + g * 32 < N)
          a.mask[(long long)s * a.mask_ld + ((row0 + g * 32) >> 5)] = word;
        if (a.pooled != nullptr && lane == 0 && word != 0u) atomicOr(a.pooled + ((row0 + g * 32) >> 5), word);
        if (valid && better(y, (int)row, bv, bi)) { bv = y; bi = (int)row; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        e.cnt[sl] += c;
        if (better(bv, bi, e.mx[sl], e.arg[sl])) { e.mx[sl] = bv; e.arg[sl] = bi; }
      }
    } else {  // EPI_TOPK
      const int k = a.k;
      float* tv = e.tv + sl * k;
      int* ti = e.ti + sl * k;
      // The list's last entry is the admission threshold: its value stays in a register and is
      // re-read after insertions.  The pre-filter is y >= threshold; ties on the value are sorted out
      // by the insertion itself (an entry that is not better than the k-th one is not inserted).
      // a.thr_seed (optional): a lower bound of the sample's k-th largest value over the whole launch
      // (from a pre-pass over the first rows); rows below it can never enter the final list
      const float seedv = a.thr_seed != nullptr ? __ldg(a.thr_seed + s) : -CUDART_INF_F;
      float thv = fmaxf(tv[k - 1], seedv);
      const bool full = (long long)row0 + TM <= N;
#pragma unroll
      for (int g = 0; g < TM / 32; ++g) {
        const int r = g * 32 + lane;
        const long long row = (long long)row0 + r;
        const float y = fmaf(sOut[sl * OUT_LD + r], a.ystd, a.c0);
        unsigned int m = __ballot_sync(0xffffffffu, (full || row < N) && y >= thv);
        if (m) {
          topk_merge_group(tv, ti, k, y, (int)row, m, lane);
          thv = fmaxf(tv[k - 1], seedv);
        }
      }
    }
  }
  }
}

// write the CTA's partial results: layout [gridDim.x][Sp] (+[k])
template <int EPI>
__device__ __forceinline__ void epilogue_flush(const KArgs& a, const EpiSmem& e, int s0, int tid,
                                               int nthreads = NTHREADS) {
  const int Sp = a.lay.Sp;
  if (EPI == EPI_LEVELSET) {
    for (int sl = tid; sl < TN; sl += nthreads) {
      const size_t o = (size_t)blockIdx.x * Sp + s0 + sl;
      a.p_cnt[o] = e.cnt[sl];
      a.p_max[o] = e.mx[sl];
      a.p_arg[o] = e.arg[sl];
    }
  } else if (EPI == EPI_TOPK) {
    const int k = a.k;
    for (int i = tid; i < TN * k; i += nthreads) {
      const int sl = i / k, q = i - sl * k;
      const size_t o = ((size_t)blockIdx.x * Sp + s0 + sl) * k + q;
      a.p_tv[o] = e.tv[i];
      a.p_ti[o] = e.ti[i];
    }
  }
}

// ---------------------------------------------------------------------------------
// feature generation for one candidate row
// ---------------------------------------------------------------------------------
// RFF feature: cos(omega . x + bias).  ft = (omega[0..DPX) | bias | ...)
template <int DP>
__device__ __forceinline__ float rff_arg(const float* __restrict__ ft, const float (&x)[DP]) {
  float arg = ft[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) arg = fmaf(x[i], ft[i], arg);
  return arg;
}

// squared scaled distance to one training point.  ft = (-Xtrain[0..DP) | ...)
template <int DP>
__device__ __forceinline__ float kern_r2(const float* __restrict__ ft, const float (&xs)[DP]) {
  float r2 = 0.f;
#pragma unroll
  for (int i = 0; i < DP; ++i) {
    const float df = xs[i] + ft[i];
    r2 = fmaf(df, df, r2);
  }
  return r2;
}

// a table row of OSV floats (multiple of 4) as 128-bit shared-memory loads
template <int OSV>
__device__ __forceinline__ void load_row(const float* __restrict__ ft, float (&w)[OSV]) {
#pragma unroll
  for (int v = 0; v < OSV / 4; ++v) {
    const float4 q = *reinterpret_cast<const float4*>(ft + 4 * v);
    w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
  }
}

// Same, with the table row fetched as 128-bit shared-memory loads (row stride and base are
// multiples of 16 bytes).
template <int DP>
__device__ __forceinline__ float rff_arg_v4(const float* __restrict__ ft, const float (&x)[DP]) {
  constexpr int NV = (DP + 1 + 3) / 4;
  float r[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = *reinterpret_cast<const float4*>(ft + 4 * v);
    r[4 * v] = q.x; r[4 * v + 1] = q.y; r[4 * v + 2] = q.z; r[4 * v + 3] = q.w;
  }
  float arg = r[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) arg = fmaf(x[i], r[i], arg);
  return arg;
}

template <int DP>
__device__ __forceinline__ float kern_r2_v4(const float* __restrict__ ft, const float (&xs)[DP]) {
  constexpr int NV = (DP + 3) / 4;
  float r[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 q = *reinterpret_cast<const float4*>(ft + 4 * v);
    r[4 * v] = q.x; r[4 * v + 1] = q.y; r[4 * v + 2] = q.z; r[4 * v + 3] = q.w;
  }
  float r2 = 0.f;
#pragma unroll
  for (int i = 0; i < DP; ++i) {
    const float df = xs[i] + r[i];
    r2 = fmaf(df, df, r2);
  }
  return r2;
}

// ---------------------------------------------------------------------------------
// Kernel A: shared basis, SIMT contraction
//   grid = (gx persistent CTAs, Sp / 64 sample chunks), block = 256
//   DP > 0: the row's x lives in DP registers; DP == 0: x staged in shared memory
// ---------------------------------------------------------------------------------
template <int DP, int EPI>
__global__ void __launch_bounds__(NTHREADS, 2) k_shared_simt(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(16) float smem[];
  const Layout& lay = a.lay;
  const int tid = threadIdx.x;
  const int OS = lay.OS, dp4 = lay.dp4, Sp = lay.Sp;
  const int s0 = blockIdx.y * TN;

  float* sState = smem;
  const EpiSmem e = epi_carve(sState, EPI, a.k);
  float* sTile = smem + round_up((int)epi_smem_floats(EPI, a.k), 4);
  float* sPhi = sTile;               // [TK][TM]
  float* sW = sPhi + TK * TM;        // [TK][TN]
  float* sFT = sW + TK * TN;         // [TK][OS]
  float* sIl = sFT + TK * OS;        // [dp4]
  float* sX = sIl + round_up(dp4, 4);  // generic only: [dp4][TM]
  float* sOut = sTile;               // alias, [TN][OUT_LD]

  const float* __restrict__ FT = a.pack + lay.off_ft;
  const float* __restrict__ WV = a.pack + lay.off_wv;
  const float* __restrict__ IL = a.pack + lay.off_invls;

  epi_init(e, EPI, a.k, tid);

  const int tx = tid & 15, ty = tid >> 4;  // contraction mapping: 4 samples x 8 rows
  const int grow = tid & (TM - 1), kh = tid >> 7;  // generation mapping
  const int nkb = lay.Kp / TK;

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    __syncthreads();  // previous tile's epilogue finished with sOut / state visible
    // ---- x for this thread's generation row ----
    float x[DP > 0 ? DP : 1], xs[DP > 0 ? DP : 1];
    {
      const long long row = (long long)row0 + grow;
      if constexpr (DP > 0) {
#pragma unroll
        for (int i = 0; i < DP; ++i) {
          x[i] = (row < a.N && i < lay.d) ? load_x(a, row, i) : 0.f;
          xs[i] = x[i] * __ldg(IL + i);
        }
      } else {
        for (int i = tid; i < dp4; i += NTHREADS) sIl[i] = __ldg(IL + i);
        const long long base = (long long)row0 * lay.d;
        const long long lim = a.N * lay.d;
        for (int idx = tid; idx < TM * lay.d; idx += NTHREADS) {
          const int r = idx / lay.d, i = idx - r * lay.d;
          sX[i * TM + r] = (base + idx < lim) ? load_x(a, (long long)row0 + r, i) : 0.f;
        }
        for (int idx = tid; idx < TM * (dp4 - lay.d); idx += NTHREADS)
          sX[(lay.d + idx / TM) * TM + (idx % TM)] = 0.f;
      }
    }
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    for (int kb = 0; kb < nkb; ++kb) {
      __syncthreads();  // previous contraction done with sPhi / sW / sFT
      {
        const float4* src = reinterpret_cast<const float4*>(FT + (size_t)kb * TK * OS);
        for (int i = tid; i < TK * OS / 4; i += NTHREADS)
          reinterpret_cast<float4*>(sFT)[i] = __ldg(src + i);
#pragma unroll
        for (int i = tid; i < TK * TN / 4; i += NTHREADS) {
          const int r = i >> 4, c4 = i & 15;
          reinterpret_cast<float4*>(sW)[i] =
              __ldg(reinterpret_cast<const float4*>(WV + (size_t)(kb * TK + r) * Sp + s0) + c4);
        }
      }
      __syncthreads();
      // ---- generate this thread's 16 features of row `grow` ----
      // Lp_ft is a multiple of 8 and every lane of a warp works on the same k, so the
      // RFF / kernel-feature branch is warp-uniform.
      {
        const int kbase = kh * (TK / 2);
#pragma unroll 4
        for (int kk = 0; kk < TK / 2; ++kk) {
          const int k = kbase + kk;
          const float* ft = sFT + k * OS;
          const bool rff = (kb * TK + k) < lay.kk0;
          const bool lin = (kb * TK + k) >= lay.Lcos;  // linear-kernel rows: identity activation
          float v;
          if constexpr (DP > 0) {
            if (rff) { const float arg = rff_arg<DP>(ft, x); v = lin ? arg : __cosf(arg); }
            else v = kernel_profile(kern_r2<DP>(ft, xs), lay.kernel);
          } else {
            if (rff) {
              float arg = ft[dp4];
              for (int i = 0; i < dp4; ++i) arg = fmaf(sX[i * TM + grow], ft[i], arg);
              v = lin ? arg : __cosf(arg);
            } else {
              float r2 = 0.f;
              for (int i = 0; i < dp4; ++i) {
                const float df = fmaf(sX[i * TM + grow], sIl[i], ft[i]);
                r2 = fmaf(df, df, r2);
              }
              v = kernel_profile(r2, lay.kernel);
            }
          }
          sPhi[k * TM + grow] = v;
        }
      }
      __syncthreads();
      // ---- contraction: acc[8 rows][4 samples] += Phi[k][rows] * W[k][samples] ----
#pragma unroll 8
      for (int k = 0; k < TK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(sPhi + k * TM + ty * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(sPhi + k * TM + ty * 8 + 4);
        const float4 b = *reinterpret_cast<const float4*>(sW + k * TN + tx * 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
      }
    }
    __syncthreads();  // everyone done reading sPhi / sW before they are overwritten
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float* dst = sOut + (tx * 4 + c) * OUT_LD + ty * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][c], acc[5][c], acc[6][c], acc[7][c]);
    }
    __syncthreads();
    epilogue_tile<EPI>(a, sOut, e, row0, s0, tid);
  }
  __syncthreads();
  epilogue_flush<EPI>(a, e, s0, tid);
}

// ---------------------------------------------------------------------------------
// Kernel B: one basis per sample (and small S).  lane = row, 4 rows per lane.
// ---------------------------------------------------------------------------------
template <int DP, int EPI>
__global__ void __launch_bounds__(NTHREADS, 2) k_persample(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(16) float smem[];
  const Layout& lay = a.lay;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int OS = lay.OS, OSB = lay.OSB, dp4 = lay.dp4, Sp = lay.Sp;
  const int s0 = blockIdx.y * TN;
  constexpr int R = TM / 32;

  float* sState = smem;
  const EpiSmem e = epi_carve(sState, EPI, a.k);
  float* sOut = smem + round_up((int)epi_smem_floats(EPI, a.k), 4);  // [TN][OUT_LD]
  float* sIl = sOut + TN * OUT_LD;                                   // [dp4]
  float* sX = sIl + round_up(dp4, 4);                                // generic: [dp4][TM]

  const float* __restrict__ FTK = a.pack + lay.off_ft + (size_t)lay.kk0 * OS;  // kernel rows
  const float* __restrict__ VT = a.pack + lay.off_wv + (size_t)lay.kk0 * Sp;   // [np][Sp]
  const float* __restrict__ PB = a.pack + lay.off_pb;
  const float* __restrict__ IL = a.pack + lay.off_invls;

  epi_init(e, EPI, a.k, tid);
  for (int i = tid; i < dp4; i += NTHREADS) sIl[i] = __ldg(IL + i);

  // work split of the chunk: ns valid samples; few samples -> split the feature range
  const int ns = min(TN, lay.S - s0);
  const int slices = (ns >= 5) ? 1 : (ns >= 3) ? 2 : (ns == 2) ? 4 : 8;

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    __syncthreads();
    float x[R][DP > 0 ? DP : 1];
    if constexpr (DP > 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long row = (long long)row0 + r * 32 + lane;
#pragma unroll
        for (int i = 0; i < DP; ++i)
          x[r][i] = (row < a.N && i < lay.d) ? load_x(a, row, i) : 0.f;
      }
    } else {
      const long long base = (long long)row0 * lay.d;
      const long long lim = a.N * lay.d;
      for (int idx = tid; idx < TM * lay.d; idx += NTHREADS) {
        const int r = idx / lay.d, i = idx - r * lay.d;
        sX[i * TM + r] = (base + idx < lim) ? load_x(a, (long long)row0 + r, i) : 0.f;
      }
      for (int idx = tid; idx < TM * (dp4 - lay.d); idx += NTHREADS)
        sX[(lay.d + idx / TM) * TM + (idx % TM)] = 0.f;
      __syncthreads();
    }

    const int nitems = (slices == 1) ? ns : ns * slices;
    for (int item = warp; item < nitems; item += NWARPS) {
      const int sl = (slices == 1) ? item : item / slices;
      const int c = (slices == 1) ? 0 : item % slices;
      const int s = s0 + sl;
      // feature ranges of this slice (multiples of 4 / 1)
      const int lq = lay.Lp / 4;
      const int l0 = 4 * ((lq * c) / slices), l1 = 4 * ((lq * (c + 1)) / slices);
      const int j0 = (lay.n * c) / slices, j1 = (lay.n * (c + 1)) / slices;
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;

      // ---- RFF part: sum_l w * cos(omega . x + b) ----
      const float* pb = PB + ((size_t)s * lay.Lp + l0) * OSB;
      if constexpr (DP == 2) {
#pragma unroll 4
        for (int l = l0; l < l1; ++l, pb += 4) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(pb));
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float arg = fmaf(x[r][0], p.x, fmaf(x[r][1], p.y, p.z));
            acc[r] = fmaf(p.w, (l >= lay.Lcos) ? arg : __cosf(arg), acc[r]);
          }
        }
      } else if constexpr (DP > 0) {
#pragma unroll 2
        for (int l = l0; l < l1; ++l, pb += OSB) {
          float om[DP];
#pragma unroll
          for (int i = 0; i < DP; i += 4) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(pb + i));
            om[i] = p.x; om[i + 1] = p.y; om[i + 2] = p.z; om[i + 3] = p.w;
          }
          const float4 bw = __ldg(reinterpret_cast<const float4*>(pb + DP));
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float arg = bw.x;
#pragma unroll
            for (int i = 0; i < DP; ++i) arg = fmaf(x[r][i], om[i], arg);
            acc[r] = fmaf(bw.y, (l >= lay.Lcos) ? arg : __cosf(arg), acc[r]);
          }
        }
      } else {
        for (int l = l0; l < l1; ++l, pb += OSB) {
          float arg[R];
          const float bb = __ldg(pb + dp4), ww = __ldg(pb + dp4 + 1);
#pragma unroll
          for (int r = 0; r < R; ++r) arg[r] = bb;
          for (int i = 0; i < dp4; i += 4) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(pb + i));
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float* xr = sX + r * 32 + lane;
              arg[r] = fmaf(xr[(i + 0) * TM], p.x, arg[r]);
              arg[r] = fmaf(xr[(i + 1) * TM], p.y, arg[r]);
              arg[r] = fmaf(xr[(i + 2) * TM], p.z, arg[r]);
              arg[r] = fmaf(xr[(i + 3) * TM], p.w, arg[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmaf(ww, (l >= lay.Lcos) ? arg[r] : __cosf(arg[r]), acc[r]);
        }
      }
      // ---- exact-kernel part: sum_j V[s,j] * kappa(|x/ls - xt_j|) ----
      for (int j = j0; j < j1; ++j) {
        const float* ft = FTK + (size_t)j * OS;
        const float vv = __ldg(VT + (size_t)j * Sp + s);
        float r2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) r2[r] = 0.f;
        if constexpr (DP > 0) {
#pragma unroll
          for (int i = 0; i < DP; ++i) {
            const float nxt = __ldg(ft + i), il = sIl[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float df = fmaf(x[r][i], il, nxt);
              r2[r] = fmaf(df, df, r2[r]);
            }
          }
        } else {
          for (int i = 0; i < dp4; ++i) {
            const float nxt = __ldg(ft + i), il = sIl[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float df = fmaf(sX[i * TM + r * 32 + lane], il, nxt);
              r2[r] = fmaf(df, df, r2[r]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(vv, kernel_profile(r2[r], lay.kernel), acc[r]);
      }
      const int orow = (slices == 1) ? sl : c * 8 + sl;
#pragma unroll
      for (int r = 0; r < R; ++r) sOut[orow * OUT_LD + r * 32 + lane] = acc[r];
    }
    __syncthreads();
    if (slices > 1) {
      // fold the feature slices in a fixed order (deterministic)
      for (int idx = tid; idx < ns * TM; idx += NTHREADS) {
        const int sl = idx / TM, r = idx - sl * TM;
        float t = sOut[sl * OUT_LD + r];
        for (int c = 1; c < slices; ++c) t += sOut[(c * 8 + sl) * OUT_LD + r];
        sOut[sl * OUT_LD + r] = t;
      }
      __syncthreads();
    }
    epilogue_tile<EPI>(a, sOut, e, row0, s0, tid);
  }
  __syncthreads();
  epilogue_flush<EPI>(a, e, s0, tid);
}

}  // namespace psbax
