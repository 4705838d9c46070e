// Feed-forward embedding of the deep-kernel GP (src/models/deep_kernel_gp.py:18-84: Linear + activation
// stack, no activation after the last layer), evaluated over all N candidates in fp32: the input
// transform of the DKGP / LinearKernel sample paths (src/utils.py:186-206, the GB1 model
// 80 -> 128 -> 64 -> 32).  One CTA = 64 candidate rows x 256 threads; the activations of a row tile
// ping-pong between two shared-memory planes [width][64], the weights of the current layer are staged
// in shared memory in column chunks ([in][OC], K-major) and every thread owns 1 row x 8 output
// columns per pass (weights by warp-broadcast LDS.128, activations conflict-free), so nothing but X*
// and the final embedding touches HBM.
#pragma once
#include <cuda_runtime.h>

#include "../../include/psbax_b200.h"

namespace psbax {

constexpr int MLP_TM = 64;         // rows per CTA
constexpr int MLP_THREADS = 256;   // 4 output groups x 64 rows
constexpr int MLP_MAXW = 256;      // widest layer (input included)
constexpr int MLP_WCHUNK = 16384;  // floats of weights staged at a time (64 KB)

struct MlpArgs {
  int n_layers, activation;
  int width[PSBAX_MLP_MAX_LAYERS + 1];
  const float* W[PSBAX_MLP_MAX_LAYERS];  // [in, out] row-major (K-major)
  const float* b[PSBAX_MLP_MAX_LAYERS];
  float slope;
  const float* X;
  float* out;
  long long N;
  int maxw;
};

__device__ __forceinline__ float mlp_act(float v, int act, float slope) {
  switch (act) {
    case PSBAX_ACT_RELU: return fmaxf(v, 0.f);
    case PSBAX_ACT_LEAKY_RELU: return v > 0.f ? v : slope * v;
    case PSBAX_ACT_SWISH: return v / (1.f + __expf(-v));
    case PSBAX_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case PSBAX_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(MLP_THREADS) k_mlp_forward(const __grid_constant__ MlpArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* act0 = smem;                               // [maxw][64]
  float* act1 = act0 + (size_t)a.maxw * MLP_TM;     // [maxw][64]
  float* sW = act1 + (size_t)a.maxw * MLP_TM;       // [in][OC]
  const int tid = threadIdx.x, r = tid & (MLP_TM - 1), og = tid >> 6;
  const long long row0 = (long long)blockIdx.x * MLP_TM;
  const int d_in = a.width[0];
  // X tile -> act0[i][r] (coalesced global reads along the row-major tile)
  for (int idx = tid; idx < MLP_TM * d_in; idx += MLP_THREADS) {
    const int rr = idx / d_in, i = idx - rr * d_in;
    act0[i * MLP_TM + rr] = (row0 + rr < a.N) ? __ldg(a.X + (row0 + rr) * d_in + i) : 0.f;
  }
  float* cur = act0;
  float* nxt = act1;
  for (int l = 0; l < a.n_layers; ++l) {
    const int in = a.width[l], out = a.width[l + 1];
    int oc = (MLP_WCHUNK / in) & ~31;  // columns per staged chunk, a multiple of 32
    if (oc > ((out + 31) & ~31)) oc = (out + 31) & ~31;
    const bool last = l == a.n_layers - 1;
    for (int o0 = 0; o0 < out; o0 += oc) {
      __syncthreads();  // previous chunk consumed / activations of the previous layer complete
      for (int idx = tid; idx < in * oc; idx += MLP_THREADS) {
        const int i = idx / oc, o = idx - i * oc;
        sW[idx] = (o0 + o < out) ? __ldg(a.W[l] + (size_t)i * out + o0 + o) : 0.f;
      }
      __syncthreads();
      for (int ob = og * 8; ob < oc; ob += 32) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const float* wp = sW + ob;
#pragma unroll 4
        for (int i = 0; i < in; ++i) {
          const float x = cur[i * MLP_TM + r];
          const float4 w0 = *reinterpret_cast<const float4*>(wp + (size_t)i * oc);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + (size_t)i * oc + 4);
          acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]);
          acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
          acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]);
          acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int o = o0 + ob + j;
          if (o < out) {
            float v = acc[j] + __ldg(a.b[l] + o);
            if (!last) v = mlp_act(v, a.activation, a.slope);
            nxt[o * MLP_TM + r] = v;
          }
        }
      }
    }
    float* t = cur; cur = nxt; nxt = t;
  }
  __syncthreads();
  const int h = a.width[a.n_layers];
  for (int idx = tid; idx < MLP_TM * h; idx += MLP_THREADS) {  // coalesced store of the [64, h] tile
    const int rr = idx / h, o = idx - rr * h;
    if (row0 + rr < a.N) a.out[(row0 + rr) * h + o] = cur[o * MLP_TM + rr];
  }
}

}  // namespace psbax
