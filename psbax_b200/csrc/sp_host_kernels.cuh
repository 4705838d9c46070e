// Small kernels launched directly by the C-ABI layer (psbax_b200.cu): partial-result reductions,
// the top-k merge and the operand pack.  Included by that translation unit only.
#pragma once
#include "sp_common.cuh"

namespace psbax {

// ---------------------------------------------------------------------------------
// finalize kernels: reduce the per-CTA partials
// ---------------------------------------------------------------------------------
// out[i] = sum over sample chunks (ascending) of part[c][i]
__global__ void k_sumsq_finalize(const float* __restrict__ part, int chunks, long long N, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += part[(long long)c * N + i];
    out[i] = acc;
  }
}

__global__ void k_levelset_finalize(const unsigned int* __restrict__ p_cnt,
                                    const float* __restrict__ p_max,
                                    const int* __restrict__ p_arg, int parts, int Sp, int S,
                                    long long row_offset, long long* __restrict__ count,
                                    float* __restrict__ maxval, long long* __restrict__ argmax) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  unsigned long long c = 0;
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int p = 0; p < parts; ++p) {
    const size_t o = (size_t)p * Sp + s;
    c += p_cnt[o];
    const float v = p_max[o];
    const int i = p_arg[o];
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
  if (count) count[s] = (long long)c;
  if (maxval) maxval[s] = bv;
  if (argmax) argmax[s] = (bi == 0x7fffffff) ? -1 : row_offset + bi;
}

// One block per sample.  Picks the k best of lists*k entries by k rounds of
// "best entry strictly after the previous pick" (deterministic, no sorting
// network, no shared-memory size limit).  in layout: [lists][ldS][k].
template <typename I>
__global__ void __launch_bounds__(128) k_topk_merge(const float* __restrict__ in_val,
                                                    const I* __restrict__ in_idx, int lists,
                                                    int ldS, int k, long long row_offset,
                                                    float* __restrict__ out_val,
                                                    long long* __restrict__ out_idx) {
  __shared__ float sv[4];
  __shared__ long long si[4];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const I SENT = (sizeof(I) == 4) ? (I)0x7fffffff : (I)0x7fffffffffffffffLL;
  float pv = CUDART_INF_F;
  long long pi = -1;
  const int total = lists * k;
  for (int round = 0; round < k; ++round) {
    float bv = -CUDART_INF_F;
    long long bi = 0x7fffffffffffffffLL;
    bool found = false;
    for (int t = tid; t < total; t += blockDim.x) {
      const int li = t / k, q = t - li * k;
      const size_t o = ((size_t)li * ldS + s) * k + q;
      const I ii = in_idx[o];
      if (ii == SENT || ii < 0) continue;
      const float v = in_val[o];
      const long long i = (long long)ii;
      const bool after = (v < pv) || (v == pv && i > pi);
      if (after && (!found || better(v, i, bv, bi))) { bv = v; bi = i; found = true; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    bv = sv[0]; bi = si[0];
#pragma unroll
    for (int w = 1; w < 4; ++w)
      if (better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
    __syncthreads();
    const bool any = bi != 0x7fffffffffffffffLL;
    if (tid == 0) {
      out_val[(size_t)s * k + round] = any ? bv : -CUDART_INF_F;
      out_idx[(size_t)s * k + round] = any ? row_offset + bi : -1;
    }
    if (!any) {
      // nothing left: fill the remaining slots
      for (int r2 = round + 1 + tid; r2 < k; r2 += blockDim.x) {
        out_val[(size_t)s * k + r2] = -CUDART_INF_F;
        out_idx[(size_t)s * k + r2] = -1;
      }
      break;
    }
    pv = bv;
    pi = bi;
  }
}

// Seed thresholds of the fused top-k: thr[s] = a value <= the k-th largest of vals[0..M)[s]
// (the k-th largest DISTINCT value, minus a relative margin of 2^-20), or -inf when there are fewer
// than k distinct values.  One block of 256 threads per sample, M <= 2048: k rounds of block max.
constexpr int TOPK_SEED_ROWS = 2048;
__global__ void __launch_bounds__(256) k_topk_seed(const float* __restrict__ vals, int M, long long ld, int k,
                                                   float* __restrict__ thr) {
  __shared__ float sw[8];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[TOPK_SEED_ROWS / 256];
#pragma unroll
  for (int j = 0; j < TOPK_SEED_ROWS / 256; ++j) {
    const int idx = tid + j * 256;
    v[j] = idx < M ? vals[(long long)idx * ld + s] : -CUDART_INF_F;
  }
  float kth = -CUDART_INF_F;
  for (int round = 0; round < k; ++round) {
    float m = v[0];
#pragma unroll
    for (int j = 1; j < TOPK_SEED_ROWS / 256; ++j) m = fmaxf(m, v[j]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) sw[warp] = m;
    __syncthreads();
    m = sw[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sw[w]);
    __syncthreads();
    kth = m;
    if (m == -CUDART_INF_F) break;
#pragma unroll
    for (int j = 0; j < TOPK_SEED_ROWS / 256; ++j)
      if (v[j] == m) v[j] = -CUDART_INF_F;  // every copy goes: the k-th distinct value is still a lower bound
  }
  if (tid == 0) thr[s] = (kth == -CUDART_INF_F) ? kth : kth - fabsf(kth) * 9.5367431640625e-7f - 1e-37f;
}

// ---------------------------------------------------------------------------------
// pack kernel
// ---------------------------------------------------------------------------------
__global__ void k_pack(Layout lay, const float* __restrict__ omega, const float* __restrict__ bias,
                       const float* __restrict__ W, const float* __restrict__ Xtrain,
                       const float* __restrict__ inv_ls, const float* __restrict__ V,
                       float* __restrict__ pack) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = lay.d, L = lay.L, n = lay.n, S = lay.S, dp4 = lay.dp4;
  const int Lcos = lay.Lcos, lin_n = lay.lin_n;  // linear-kernel rows: omega := Xtrain * inv_ls, w := V
  // invls
  for (size_t i = t0; i < (size_t)round_up(dp4, 4); i += stride)
    pack[lay.off_invls + i] = (i < (size_t)d && inv_ls) ? inv_ls[i] : 0.f;
  // ft
  {
    const int OS = lay.OS, Lpf = lay.Lp_ft, kk0 = lay.kk0;  // rows in [Lpf, kk0): zero padding
    const size_t tot = (size_t)lay.Kp * OS;
    for (size_t idx = t0; idx < tot; idx += stride) {
      const int row = (int)(idx / OS), c = (int)(idx - (size_t)row * OS);
      float v = 0.f;
      if (row < kk0) {
        if (row >= Lpf) v = 0.f;
        else if (row < L) v = (c < d) ? omega[(size_t)row * d + c] : (c == dp4 ? bias[row] : 0.f);
        else if (row >= Lcos && row - Lcos < lin_n && c < d) v = Xtrain[(size_t)(row - Lcos) * d + c] * inv_ls[c];
      } else {
        const int j = row - kk0;
        if (j < n && c < d) v = -Xtrain[(size_t)j * d + c];
      }
      pack[lay.off_ft + idx] = v;
    }
  }
  // wv
  {
    const int Sp = lay.Sp, Lpf = lay.Lp_ft, kk0 = lay.kk0;
    const size_t tot = (size_t)lay.Kp * Sp;
    for (size_t idx = t0; idx < tot; idx += stride) {
      const int row = (int)(idx / Sp), s = (int)(idx - (size_t)row * Sp);
      float v = 0.f;
      if (s < S) {
        if (row < kk0) {
          if (row >= Lpf) v = 0.f;
          else if (row < L) v = W[(size_t)s * L + row];
          else if (row >= Lcos && row - Lcos < lin_n) v = V[(size_t)s * lin_n + (row - Lcos)];
        } else {
          const int j = row - kk0;
          if (j < n) v = V[(size_t)s * n + j];
        }
      }
      pack[lay.off_wv + idx] = v;
    }
  }
  // pb
  if (lay.mode == MODE_PERSAMPLE) {
    const int OSB = lay.OSB, Lp = lay.Lp;
    const size_t tot = (size_t)S * Lp * OSB;
    for (size_t idx = t0; idx < tot; idx += stride) {
      const int c = (int)(idx % OSB);
      const size_t sl = idx / OSB;
      const int l = (int)(sl % Lp), s = (int)(sl / Lp);
      const int g = (lay.G == 1) ? 0 : s;
      float v = 0.f;
      if (l < L) {
        if (c < d) v = omega[((size_t)g * L + l) * d + c];
        else if (c == dp4) v = bias[(size_t)g * L + l];
        else if (c == dp4 + 1) v = W[(size_t)s * L + l];
      } else if (l >= Lcos && l - Lcos < lin_n) {
        if (c < d) v = Xtrain[(size_t)(l - Lcos) * d + c] * inv_ls[c];
        else if (c == dp4 + 1) v = V[(size_t)s * lin_n + (l - Lcos)];
      }
      pack[lay.off_pb + idx] = v;
    }
  }
}

}  // namespace psbax
