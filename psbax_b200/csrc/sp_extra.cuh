// Secondary kernels on the path: DiscoBAX subset selection and sample-path gradients.
#pragma once
#include "sp_common.cuh"

namespace psbax {

// ---------------------------------------------------------------------------------
// SubsetSelect (src/algorithms/discobax.py:37-56).  One CTA works on T samples at
// once so that every eta[b, j] fetched from L2 is used T times.
//   values[b, j] = fx[j, s] + etas[b, j]   (max(0, .) when nonneg)
//   pick 0      = arg max_j mean_b values[b, j]
//   pick t >= 1 = arg max_j mean_b max(cur[b], values[b, j]), chosen j score 0
// ---------------------------------------------------------------------------------
// Register blocking: a thread scores SS_J candidates at once and keeps the running maxima of
// four noise draws x T samples in registers, so the inner (draw, candidate, sample) step is
// three ALU instructions (add, max, add) with one eta load per (draw, candidate) and one
// 16-byte shared-memory load per (four draws, sample).  cur starts at -inf, so the first pick
// needs no special case; sums run over b in ascending order (same rounding as a scalar loop).
constexpr int SS_J = 2;
// packed FP32 (two samples per 64-bit register pair): the two adds of a step become one FADD2 each
// on the FMA pipe, the max stays scalar on the ALU pipe -> the pipes are balanced at 2 issue
// cycles per (draw, candidate, sample) instead of 4 on the FMA pipe.  Same IEEE operations.
__device__ __forceinline__ uint64_t ss_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ss_unpack2(uint64_t r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ uint64_t ss_fadd2(uint64_t a, uint64_t b) {
  uint64_t c;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  return c;
}
template <int T, bool NONNEG>
__global__ void __launch_bounds__(512) k_subset_select(const float* __restrict__ fx,
                                                      long long ld_fx,
                                                      const float* __restrict__ etas,
                                                      long long N, int B, int S, int k,
                                                      long long* __restrict__ out_idx) {
  extern __shared__ __align__(16) float smem[];
  const int Bp = (B + 3) & ~3;
  float* cur = smem;                                    // [T][Bp]
  int* chosen = reinterpret_cast<int*>(cur + T * Bp);   // [T][k]
  float* redv = reinterpret_cast<float*>(chosen + T * k);  // [T][32]
  int* redi = reinterpret_cast<int*>(redv + T * 32);       // [T][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sbase = blockIdx.x * T;
  const float invB = 1.0f / (float)B;
  const int B4 = B & ~3;
  for (int i = tid; i < T * Bp; i += blockDim.x) cur[i] = -CUDART_INF_F;
  __syncthreads();

  for (int step = 0; step < k; ++step) {
    float bv[T];
    int bi[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { bv[t] = -CUDART_INF_F; bi[t] = 0x7fffffff; }
    for (long long j0 = tid; j0 < N; j0 += (long long)SS_J * blockDim.x) {
      float f[SS_J][T], acc[SS_J][T];
      long long jr[SS_J];
#pragma unroll
      for (int jj = 0; jj < SS_J; ++jj) {
        const long long j = j0 + (long long)jj * blockDim.x;
        jr[jj] = j < N ? j : j0;  // out-of-range slot repeats row j0 (never scored)
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int s = sbase + t;
          f[jj][t] = (s < S) ? __ldg(fx + jr[jj] * ld_fx + s) : 0.f;
          acc[jj][t] = 0.f;
        }
      }
      if constexpr (T >= 2) {
        constexpr int TP = T / 2;
        uint64_t f2[SS_J][TP], acc2[SS_J][TP];
#pragma unroll
        for (int jj = 0; jj < SS_J; ++jj)
#pragma unroll
          for (int tp = 0; tp < TP; ++tp) {
            f2[jj][tp] = ss_pack2(f[jj][2 * tp], f[jj][2 * tp + 1]);
            acc2[jj][tp] = 0ull;
          }
        // the eta values of the next four draws are fetched before this chunk's arithmetic (four
        // warps per scheduler cannot hide an L2 round trip on their own)
        float eta_n[4][SS_J];
        if (B4 > 0) {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int jj = 0; jj < SS_J; ++jj) eta_n[bb][jj] = __ldg(etas + (long long)bb * N + jr[jj]);
        }
        for (int b0 = 0; b0 < B4; b0 += 4) {
          float eta_c[4][SS_J];
          const int bn = (b0 + 4 < B4) ? b0 + 4 : b0;  // last chunk re-reads itself (values unused)
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int jj = 0; jj < SS_J; ++jj) {
              eta_c[bb][jj] = eta_n[bb][jj];
              eta_n[bb][jj] = __ldg(etas + (long long)(bn + bb) * N + jr[jj]);
            }
          float c[T][4];
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const float4 c4 = *reinterpret_cast<const float4*>(cur + t * Bp + b0);
            c[t][0] = c4.x; c[t][1] = c4.y; c[t][2] = c4.z; c[t][3] = c4.w;
          }
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int jj = 0; jj < SS_J; ++jj) {
              const uint64_t eta2 = ss_pack2(eta_c[bb][jj], eta_c[bb][jj]);
#pragma unroll
              for (int tp = 0; tp < TP; ++tp) {
                float lo, hi;
                ss_unpack2(ss_fadd2(f2[jj][tp], eta2), lo, hi);
                if (NONNEG) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                acc2[jj][tp] = ss_fadd2(acc2[jj][tp], ss_pack2(fmaxf(lo, c[2 * tp][bb]), fmaxf(hi, c[2 * tp + 1][bb])));
              }
              if constexpr (T & 1) {  // odd T: the last sample is not part of a pair
                float v = f[jj][T - 1] + eta_c[bb][jj];
                if (NONNEG) v = fmaxf(v, 0.f);
                acc[jj][T - 1] += fmaxf(v, c[T - 1][bb]);
              }
            }
        }
#pragma unroll
        for (int jj = 0; jj < SS_J; ++jj)
#pragma unroll
          for (int tp = 0; tp < TP; ++tp) ss_unpack2(acc2[jj][tp], acc[jj][2 * tp], acc[jj][2 * tp + 1]);
      } else {
        for (int b0 = 0; b0 < B4; b0 += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(cur + b0);
          const float c[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int jj = 0; jj < SS_J; ++jj) {
              float v = f[jj][0] + __ldg(etas + (long long)(b0 + bb) * N + jr[jj]);
              if (NONNEG) v = fmaxf(v, 0.f);
              acc[jj][0] += fmaxf(v, c[bb]);
            }
        }
      }
      for (int b = B4; b < B; ++b)
#pragma unroll
        for (int jj = 0; jj < SS_J; ++jj) {
          const float eta = __ldg(etas + (long long)b * N + jr[jj]);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            float v = f[jj][t] + eta;
            if (NONNEG) v = fmaxf(v, 0.f);
            acc[jj][t] += fmaxf(v, cur[t * Bp + b]);
          }
        }
#pragma unroll
      for (int jj = 0; jj < SS_J; ++jj) {
        const long long j = j0 + (long long)jj * blockDim.x;
        if (j < N) {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            float score = acc[jj][t] * invB;
            for (int q = 0; q < step; ++q)
              if (chosen[t * k + q] == (int)j) score = 0.f;
            if (better(score, (int)j, bv[t], bi[t])) { bv[t] = score; bi[t] = (int)j; }
          }
        }
      }
    }
    // block arg max per sample
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv[t], off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi[t], off);
        if (better(ov, oi, bv[t], bi[t])) { bv[t] = ov; bi[t] = oi; }
      }
      if (lane == 0) { redv[t * 32 + warp] = bv[t]; redi[t * 32 + warp] = bi[t]; }
    }
    __syncthreads();
    if (tid < T) {
      float v = redv[tid * 32];
      int i = redi[tid * 32];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (better(redv[tid * 32 + w], redi[tid * 32 + w], v, i)) { v = redv[tid * 32 + w]; i = redi[tid * 32 + w]; }
      chosen[tid * k + step] = i;
      if (sbase + tid < S) out_idx[(long long)(sbase + tid) * k + step] = i;
    }
    __syncthreads();
    // cur[b] = max(cur[b], values[b, pick])
    for (int idx = tid; idx < T * B; idx += blockDim.x) {
      const int t = idx / B, b = idx - t * B;
      const int s = sbase + t;
      if (s < S) {
        const int j = chosen[t * k + step];
        float v = __ldg(fx + (long long)j * ld_fx + s) + __ldg(etas + (long long)b * N + j);
        if (NONNEG) v = fmaxf(v, 0.f);
        cur[t * Bp + b] = fmaxf(cur[t * Bp + b], v);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------
// SubsetSelect, general noise model (src/problems.py:146-163, get_noisy_f_lst): one CTA per sample,
//   values[b, j] = fx[j, s] + etas[b, j]   (MUL == false)   or   fx[j, s] * etas[b, j]   (MUL == true: the
//   multiplicative model, etas = the Bernoulli masks the objective draws, 0 / 1),
// optionally max(0, .), with one noise matrix for all samples (eta_stride_s == 0) or one per sample
// (the reference draws a fresh mask in every get_noisy_f_lst call, i.e. per sample path).  Same
// greedy, same tie rule (smaller index) as k_subset_select; the additive shared-noise case uses the
// register-blocked kernel above.
// ---------------------------------------------------------------------------------
template <bool MUL, bool NONNEG>
__global__ void __launch_bounds__(256) k_subset_select_general(const float* __restrict__ fx, long long ld_fx,
                                                               const float* __restrict__ etas,
                                                               long long eta_stride_s, long long N, int B, int k,
                                                               long long* __restrict__ out_idx) {
  extern __shared__ __align__(16) float smem[];
  float* cur = smem;                                   // [B]
  int* chosen = reinterpret_cast<int*>(cur + B);       // [k]
  __shared__ float redv[8];
  __shared__ int redi[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x;
  const float* __restrict__ E = etas + (long long)s * eta_stride_s;
  const float invB = 1.0f / (float)B;
  for (int b = tid; b < B; b += blockDim.x) cur[b] = -CUDART_INF_F;
  __syncthreads();
  for (int step = 0; step < k; ++step) {
    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    for (long long j = tid; j < N; j += blockDim.x) {
      const float f = __ldg(fx + j * ld_fx + s);
      float acc = 0.f;
      for (int b = 0; b < B; ++b) {  // ascending b: same rounding as the host loop
        const float e = __ldg(E + (long long)b * N + j);
        float v = MUL ? f * e : f + e;
        if (NONNEG) v = fmaxf(v, 0.f);
        acc += fmaxf(v, cur[b]);
      }
      float score = acc * invB;
      for (int q = 0; q < step; ++q)
        if (chosen[q] == (int)j) score = 0.f;
      if (better(score, (int)j, bv, bi)) { bv = score; bi = (int)j; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { redv[warp] = bv; redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = redv[0];
      int i = redi[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (better(redv[w], redi[w], v, i)) { v = redv[w]; i = redi[w]; }
      chosen[step] = i;
      out_idx[(long long)s * k + step] = i;
    }
    __syncthreads();
    const int jp = chosen[step];
    const float fp = __ldg(fx + (long long)jp * ld_fx + s);
    for (int b = tid; b < B; b += blockDim.x) {
      const float e = __ldg(E + (long long)b * N + jp);
      float v = MUL ? fp * e : fp + e;
      if (NONNEG) v = fmaxf(v, 0.f);
      cur[b] = fmaxf(cur[b], v);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------
// Gradient of one sample path per point (LBFGSB, src/algorithms/lbfgsb.py:29-39).
// One CTA (128 threads) per point: phase 1 stores coef_l = -w_l sin(arg_l) (and the
// kernel-term coefficients) in shared memory and reduces the value; phase 2 thread
// i < d accumulates grad[i] = sum_l coef_l * omega[l][i] (+ kernel term).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_eval_grad(const __grid_constant__ KArgs a,
                                                   const int* __restrict__ sample_idx,
                                                   float* __restrict__ val,
                                                   float* __restrict__ grad) {
  extern __shared__ __align__(16) float smem[];
  const Layout& lay = a.lay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long p = blockIdx.x;
  const int s = sample_idx ? sample_idx[p] : 0;
  const int d = lay.d, dp4 = lay.dp4;
  float* sx = smem;                 // [dp4]
  float* sxs = sx + dp4;            // [dp4]  x * inv_ls
  float* coef = sxs + dp4;          // [Lp]
  float* coefk = coef + lay.Lp;     // [np]
  __shared__ float red[4];

  const float* __restrict__ IL = a.pack + lay.off_invls;
  for (int i = tid; i < dp4; i += blockDim.x) {
    const float xv = (i < d) ? __ldg(a.X + p * d + i) : 0.f;
    sx[i] = xv;
    sxs[i] = xv * __ldg(IL + i);
  }
  __syncthreads();

  const bool persample = lay.mode == MODE_PERSAMPLE;
  const float* __restrict__ rows = persample ? a.pack + lay.off_pb + (size_t)s * lay.Lp * lay.OSB
                                             : a.pack + lay.off_ft;
  const int rs = persample ? lay.OSB : lay.OS;
  const float* __restrict__ WV = a.pack + lay.off_wv;
  float acc = 0.f;
  for (int l = tid; l < lay.Lp; l += blockDim.x) {
    const float* r = rows + (size_t)l * rs;
    float arg = __ldg(r + dp4);
    for (int i = 0; i < d; ++i) arg = fmaf(sx[i], __ldg(r + i), arg);
    const float w = persample ? __ldg(r + dp4 + 1) : __ldg(WV + (size_t)l * lay.Sp + s);
    if (l >= lay.Lcos) {  // linear-kernel row: value w * arg, gradient w * omega
      acc = fmaf(w, arg, acc);
      coef[l] = w;
    } else {
      float sn, cs;
      sincosf(arg, &sn, &cs);
      acc = fmaf(w, cs, acc);
      coef[l] = -w * sn;
    }
  }
  const float* __restrict__ FTK = a.pack + lay.off_ft + (size_t)lay.kk0 * lay.OS;
  const float* __restrict__ VT = a.pack + lay.off_wv + (size_t)lay.kk0 * lay.Sp;
  for (int j = tid; j < lay.n; j += blockDim.x) {
    const float* r = FTK + (size_t)j * lay.OS;
    float r2 = 0.f;
    for (int i = 0; i < d; ++i) {
      const float df = sxs[i] + __ldg(r + i);
      r2 = fmaf(df, df, r2);
    }
    const float v = __ldg(VT + (size_t)j * lay.Sp + s);
    acc = fmaf(v, kernel_profile(r2, lay.kernel), acc);
    coefk[j] = v * kernel_profile_dr2x2(r2, lay.kernel);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0 && val) val[p] = fmaf(red[0] + red[1] + red[2] + red[3], a.ystd, a.c0);
  for (int i = tid; i < d; i += blockDim.x) {
    float g = 0.f;
    for (int l = 0; l < lay.Lp; ++l) g = fmaf(coef[l], __ldg(rows + (size_t)l * rs + i), g);
    const float il = __ldg(IL + i);
    float gk = 0.f;
    for (int j = 0; j < lay.n; ++j) gk = fmaf(coefk[j], sxs[i] + __ldg(FTK + (size_t)j * lay.OS + i), gk);
    grad[p * d + i] = a.ystd * (g + gk * il);
  }
}

// ---------------------------------------------------------------------------------
// Batched multi-start ascent (LBFGSB, src/algorithms/lbfgsb.py:24-40 -> botorch optimize_acqf:
// num_restarts L-BFGS-B ascents from the best raw samples, box [0, 1]^d, maxiter 100).
// One CTA (128 threads) per (start, sample) pair runs the WHOLE ascent on the device: value + gradient
// of the path by all threads (as k_eval_grad), then a projected L-BFGS step (two-loop recursion over
// the last ASC_MEM curvature pairs on the free variables, Armijo backtracking along the projected
// path) by thread 0 — the vector algebra is d <= 32 wide, the evaluation is L + n wide.  Every pair
// is independent: restarts x samples CTAs in one launch instead of one launch per iterate.
// ---------------------------------------------------------------------------------
constexpr int ASC_MEM = 8;
constexpr int ASC_DMAX = 32;

struct AscShared {
  float x[ASC_DMAX], xt[ASC_DMAX], g[ASC_DMAX], gt[ASC_DMAX], p[ASC_DMAX];
  float sk[ASC_MEM][ASC_DMAX], yk[ASC_MEM][ASC_DMAX], rho[ASC_MEM];
  float f, ft, red[4];
  int flag;
};

// f_s(x) (output units) and d f_s / dx at the point in `xin` (shared, [d]); result -> *fout, gout[d]
// (shared).  Called by all 128 threads; ends with a barrier.
__device__ __forceinline__ void asc_value_grad(const KArgs& a, int s, const float* xin, float* sx, float* sxs,
                                               float* coef, float* coefk, float* red, float* fout, float* gout) {
  const Layout& lay = a.lay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = lay.d, dp4 = lay.dp4;
  const float* __restrict__ IL = a.pack + lay.off_invls;
  for (int i = tid; i < dp4; i += blockDim.x) {
    const float xv = (i < d) ? xin[i] : 0.f;
    sx[i] = xv;
    sxs[i] = xv * __ldg(IL + i);
  }
  __syncthreads();
  const bool persample = lay.mode == MODE_PERSAMPLE;
  const float* __restrict__ rows = persample ? a.pack + lay.off_pb + (size_t)s * lay.Lp * lay.OSB
                                             : a.pack + lay.off_ft;
  const int rs = persample ? lay.OSB : lay.OS;
  const float* __restrict__ WV = a.pack + lay.off_wv;
  float acc = 0.f;
  for (int l = tid; l < lay.Lp; l += blockDim.x) {
    const float* r = rows + (size_t)l * rs;
    float arg = __ldg(r + dp4);
    for (int i = 0; i < d; ++i) arg = fmaf(sx[i], __ldg(r + i), arg);
    const float w = persample ? __ldg(r + dp4 + 1) : __ldg(WV + (size_t)l * lay.Sp + s);
    if (l >= lay.Lcos) {
      acc = fmaf(w, arg, acc);
      coef[l] = w;
    } else {
      float sn, cs;
      sincosf(arg, &sn, &cs);
      acc = fmaf(w, cs, acc);
      coef[l] = -w * sn;
    }
  }
  const float* __restrict__ FTK = a.pack + lay.off_ft + (size_t)lay.kk0 * lay.OS;
  const float* __restrict__ VT = a.pack + lay.off_wv + (size_t)lay.kk0 * lay.Sp;
  for (int j = tid; j < lay.n; j += blockDim.x) {
    const float* r = FTK + (size_t)j * lay.OS;
    float r2 = 0.f;
    for (int i = 0; i < d; ++i) {
      const float df = sxs[i] + __ldg(r + i);
      r2 = fmaf(df, df, r2);
    }
    const float v = __ldg(VT + (size_t)j * lay.Sp + s);
    acc = fmaf(v, kernel_profile(r2, lay.kernel), acc);
    coefk[j] = v * kernel_profile_dr2x2(r2, lay.kernel);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) *fout = fmaf(red[0] + red[1] + red[2] + red[3], a.ystd, a.c0);
  // gradient: warp w takes dimensions w, w + 4, ...; lanes split the feature range
  for (int i = warp; i < d; i += 4) {
    float g = 0.f;
    for (int l = lane; l < lay.Lp; l += 32) g = fmaf(coef[l], __ldg(rows + (size_t)l * rs + i), g);
    float gk = 0.f;
    for (int j = lane; j < lay.n; j += 32) gk = fmaf(coefk[j], sxs[i] + __ldg(FTK + (size_t)j * lay.OS + i), gk);
    g = fmaf(gk, __ldg(IL + i), g);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
    if (lane == 0) gout[i] = a.ystd * g;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) k_ascend(const __grid_constant__ KArgs a, const int* __restrict__ sample_idx,
                                                float* __restrict__ xout, float* __restrict__ fout,
                                                int* __restrict__ iters, int maxiter, float pgtol, float ftol) {
  extern __shared__ __align__(16) float smem[];
  __shared__ AscShared sh;
  const Layout& lay = a.lay;
  const int tid = threadIdx.x;
  const long long p = blockIdx.x;
  const int s = sample_idx ? sample_idx[p] : 0;
  const int d = lay.d, dp4 = lay.dp4;
  float* sx = smem;
  float* sxs = sx + dp4;
  float* coef = sxs + dp4;
  float* coefk = coef + lay.Lp;
  if (tid < d) sh.x[tid] = fminf(fmaxf(__ldg(a.X + p * d + tid), 0.f), 1.f);
  __syncthreads();
  asc_value_grad(a, s, sh.x, sx, sxs, coef, coefk, sh.red, &sh.f, sh.g);
  int nmem = 0, head = 0, it = 0;  // thread 0's view of the curvature history (ring)
  for (; it < maxiter; ++it) {
    if (tid == 0) {
      // ascent on f == descent on h = -f; free variables: not pinned at a bound by the gradient
      float q[ASC_DMAX], al[ASC_MEM];
      float pg = 0.f;
      for (int i = 0; i < d; ++i) {
        const float gi = sh.g[i];
        const bool pinned = (sh.x[i] <= 0.f && gi < 0.f) || (sh.x[i] >= 1.f && gi > 0.f);
        q[i] = pinned ? 0.f : -gi;  // gradient of h on the free set
        pg = fmaxf(pg, fabsf(q[i]));
      }
      int flag = 0;
      if (pg <= pgtol) flag = 1;  // projected gradient small: converged
      if (!flag) {
        for (int m = 0; m < nmem; ++m) {  // two-loop recursion, newest pair first
          const int j = (head - 1 - m + ASC_MEM) % ASC_MEM;
          float dt = 0.f;
          for (int i = 0; i < d; ++i) dt = fmaf(sh.sk[j][i], q[i], dt);
          al[m] = sh.rho[j] * dt;
          for (int i = 0; i < d; ++i) q[i] = fmaf(-al[m], sh.yk[j][i], q[i]);
        }
        if (nmem > 0) {
          const int j = (head - 1 + ASC_MEM) % ASC_MEM;
          float yy = 0.f;
          for (int i = 0; i < d; ++i) yy = fmaf(sh.yk[j][i], sh.yk[j][i], yy);
          const float gamma = 1.f / (sh.rho[j] * yy);
          for (int i = 0; i < d; ++i) q[i] *= gamma;
        }
        for (int m = nmem - 1; m >= 0; --m) {
          const int j = (head - 1 - m + ASC_MEM) % ASC_MEM;
          float dt = 0.f;
          for (int i = 0; i < d; ++i) dt = fmaf(sh.yk[j][i], q[i], dt);
          const float be = sh.rho[j] * dt;
          for (int i = 0; i < d; ++i) q[i] = fmaf(al[m] - be, sh.sk[j][i], q[i]);
        }
        // p = -H grad h on the free set; fall back to steepest ascent when it is not an ascent direction of f
        float slope = 0.f, gn = 0.f;
        for (int i = 0; i < d; ++i) {
          const float gi = sh.g[i];
          const bool pinned = (sh.x[i] <= 0.f && gi < 0.f) || (sh.x[i] >= 1.f && gi > 0.f);
          sh.p[i] = pinned ? 0.f : -q[i];
          slope = fmaf(sh.p[i], gi, slope);
          gn = fmaf(pinned ? 0.f : gi, gi, gn);
        }
        if (!(slope > 1e-12f * gn)) {
          nmem = 0;  // drop the history, restart from a unit-length steepest-ascent step
          head = 0;
          const float sc = 1.f / fmaxf(sqrtf(gn), 1e-12f);
          for (int i = 0; i < d; ++i) {
            const float gi = sh.g[i];
            const bool pinned = (sh.x[i] <= 0.f && gi < 0.f) || (sh.x[i] >= 1.f && gi > 0.f);
            sh.p[i] = pinned ? 0.f : sc * gi;
          }
        } else if (nmem == 0) {  // first step: unit length (scipy's L-BFGS-B starts with 1 / |g|)
          const float sc = 1.f / fmaxf(sqrtf(gn), 1e-12f);
          for (int i = 0; i < d; ++i) sh.p[i] *= sc;
        }
      }
      sh.flag = flag;
    }
    __syncthreads();
    if (sh.flag) break;
    // backtracking along the projected path x(t) = clip(x + t p)
    float t = 1.f;
    bool accepted = false;
    for (int ls = 0; ls < 20; ++ls) {
      if (tid < d) sh.xt[tid] = fminf(fmaxf(fmaf(t, sh.p[tid], sh.x[tid]), 0.f), 1.f);
      __syncthreads();
      asc_value_grad(a, s, sh.xt, sx, sxs, coef, coefk, sh.red, &sh.ft, sh.gt);
      if (tid == 0) {
        float lin = 0.f;
        for (int i = 0; i < d; ++i) lin = fmaf(sh.g[i], sh.xt[i] - sh.x[i], lin);
        sh.flag = (sh.ft >= sh.f + 1e-4f * lin && lin > 0.f) ? 1 : 0;
      }
      __syncthreads();
      if (sh.flag) { accepted = true; break; }
      __syncthreads();
      t *= 0.5f;
    }
    if (!accepted) break;  // no progress along the direction: stop at the current point
    if (tid == 0) {
      float sy = 0.f, yy = 0.f;
      const int j = head;
      for (int i = 0; i < d; ++i) {
        const float si = sh.xt[i] - sh.x[i];
        const float yi = -(sh.gt[i] - sh.g[i]);  // gradient difference of h = -f
        sh.sk[j][i] = si;
        sh.yk[j][i] = yi;
        sy = fmaf(si, yi, sy);
        yy = fmaf(yi, yi, yy);
      }
      if (sy > 1e-10f * yy && yy > 0.f) {
        sh.rho[j] = 1.f / sy;
        head = (head + 1) % ASC_MEM;
        nmem = min(nmem + 1, ASC_MEM);
      }
      const float df = sh.ft - sh.f;
      sh.flag = (df <= ftol * fmaxf(fmaxf(fabsf(sh.f), fabsf(sh.ft)), 1.f)) ? 1 : 0;
      for (int i = 0; i < d; ++i) {
        sh.x[i] = sh.xt[i];
        sh.g[i] = sh.gt[i];
      }
      sh.f = sh.ft;
    }
    __syncthreads();
    if (sh.flag) { ++it; break; }
  }
  if (tid < d) xout[p * d + tid] = sh.x[tid];
  if (tid == 0) {
    fout[p] = sh.f;
    if (iters) iters[p] = it;
  }
}

}  // namespace psbax
