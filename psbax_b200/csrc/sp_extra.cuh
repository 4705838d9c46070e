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
template <int T>
__global__ void __launch_bounds__(1024) k_subset_select(const float* __restrict__ fx,
                                                       long long ld_fx,
                                                       const float* __restrict__ etas,
                                                       long long N, int B, int S, int k,
                                                       int nonneg, long long* __restrict__ out_idx) {
  extern __shared__ __align__(16) float smem[];
  float* cur = smem;                                   // [T][B]
  int* chosen = reinterpret_cast<int*>(cur + T * B);   // [T][k]
  float* redv = reinterpret_cast<float*>(chosen + T * k);  // [T][32]
  int* redi = reinterpret_cast<int*>(redv + T * 32);       // [T][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sbase = blockIdx.x * T;
  const float invB = 1.0f / (float)B;

  for (int step = 0; step < k; ++step) {
    float bv[T];
    int bi[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { bv[t] = -CUDART_INF_F; bi[t] = 0x7fffffff; }
    for (long long j = tid; j < N; j += blockDim.x) {
      float f[T], acc[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int s = sbase + t;
        f[t] = (s < S) ? __ldg(fx + j * ld_fx + s) : 0.f;
        acc[t] = 0.f;
      }
      for (int b = 0; b < B; ++b) {
        const float eta = __ldg(etas + (long long)b * N + j);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          float v = f[t] + eta;
          if (nonneg) v = fmaxf(v, 0.f);
          if (step > 0) v = fmaxf(v, cur[t * B + b]);
          acc[t] += v;
        }
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float score = acc[t] * invB;
        if (step > 0) {
          for (int q = 0; q < step; ++q)
            if (chosen[t * k + q] == (int)j) score = 0.f;
        }
        if (better(score, (int)j, bv[t], bi[t])) { bv[t] = score; bi[t] = (int)j; }
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
        if (nonneg) v = fmaxf(v, 0.f);
        cur[idx] = (step > 0) ? fmaxf(cur[idx], v) : v;
      }
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
  const float* __restrict__ FTK = a.pack + lay.off_ft + (size_t)lay.Lp_ft * lay.OS;
  const float* __restrict__ VT = a.pack + lay.off_wv + (size_t)lay.Lp_ft * lay.Sp;
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

}  // namespace psbax
