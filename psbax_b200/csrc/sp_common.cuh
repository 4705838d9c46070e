// Shared definitions for the PS-BAX sample-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/psbax_b200.h"

namespace psbax {

constexpr int TM = 128;        // candidate rows per tile
constexpr int TN = 64;         // sample paths per chunk
constexpr int TK = 32;         // feature block of the SIMT contraction
constexpr int OUT_LD = TM + 4; // leading dimension of the staged [TN][TM] value tile
constexpr int KMAX = 64;       // largest supported k for the fused top-k
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;

enum { EPI_VALUES = 0, EPI_LEVELSET = 1, EPI_TOPK = 2 };
enum { MODE_PERSAMPLE = 0, MODE_SHARED_SIMT = 1, MODE_TENSOR = 2, MODE_TENSOR_BF16 = 3 };

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Device-side operand pack.  All offsets in floats from the pack base.
//   invls [dp4]                 1/lengthscale, zero padded
//   ft    [Kp][OS]              shared-basis feature table: rows < Lp are
//                               (omega[0..dp4) | bias | 0..), rows Lp.. are
//                               (-Xtrain[0..dp4) | 0..).  Per-sample mode keeps
//                               only the kernel rows ([np][OS], Lp_ft = 0).
//   wv    [Kp][Sp]              [W ; V] stacked, K-major (per-sample mode: V rows only)
//   pb    [S][Lp][OSB]          per-sample table (omega[0..dp4) | bias | w | 0..)
//   tc    tensor-engine operands (see sp_tensor.cuh)
struct Layout {
  int d, L, G, S, n, kernel;
  int dp4;    // x registers / padded dimension: 2 for d <= 2, else round_up(d, 4)
  int OS;     // row stride of ft
  int OSB;    // row stride of pb
  int Lp;     // L rounded up to 8
  int np_;    // n rounded up to 8
  int Lp_ft;  // RFF rows present in ft / wv (Lp in shared modes, 0 in per-sample mode)
  int Lcos;   // feature rows below Lcos use cos(arg); rows in [Lcos, Lp) are linear-kernel rows: arg itself
  int lin_n;  // number of linear-kernel rows (PSBAX_KERNEL_LINEAR), else 0
  int Kp;     // rows of ft / wv
  int kk0;    // first exact-kernel row of ft / wv (== Lp_ft unless the BF16x3 engine pads the RFF section
              // to whole blocks); rows in [Lp_ft, kk0) are padding (zero table rows, zero weights)
  int nkb_r;  // tensor engines: feature blocks of tkb rows covering [0, kk0) (every block when nkb_k == 0)
  int nkb_k;  // BF16x3 engine: exact-kernel blocks of tkb / 2 rows each, contracted in 3xTF32 (see below)
  int Sp;     // S rounded up to TN
  int mode;
  int tkb;         // tensor engines: features per block (32, or 64 for BF16x3 with x in registers)
  int tc_generic;  // tensor engines: x tile in shared memory + streamed table blocks
  int tc_proj;     // BF16x3 engine, projection variant: feature arguments computed by 3xTF32 MMAs (sp_tensor_proj.cuh)
  size_t off_invls, off_ft, off_wv, off_pb, off_tc, total_floats;
};

// Block structure of the feature axis.  The BF16x3 tensor engine keeps the exact-kernel rows
// k(x, X_train) . V_s out of the bf16 split: V = (K + sigma^2 I)^-1 (...) has large entries that
// cancel in the sum, and a 2^-17 operand error there is far above 1e-4 of the path magnitude.  Its
// feature axis is therefore [RFF rows, padded to whole blocks of tkb][kernel rows, blocks of tkb / 2]
// and the kernel blocks are contracted with the 3xTF32 split (same stage footprints: tkb bf16
// features and tkb / 2 tf32 features fill the same tensor-memory columns and operand bytes).
__host__ __device__ inline void layout_blocks(Layout& l) {
  const bool mixed = (l.mode == MODE_TENSOR_BF16) && l.np_ > 0;
  if (mixed) {
    const int kf = l.tkb / 2;
    const int npk = round_up(l.np_, kf);
    l.kk0 = round_up(l.Lp_ft, l.tkb);
    l.Kp = l.kk0 + npk;
    l.nkb_r = l.kk0 / l.tkb;
    l.nkb_k = npk / kf;
  } else {
    l.kk0 = l.Lp_ft;
    l.Kp = round_up(l.Lp_ft + l.np_, l.tkb);
    l.nkb_r = l.Kp / l.tkb;
    l.nkb_k = 0;
  }
}

struct KArgs {
  Layout lay;
  const float* pack;
  const float* X;
  long long N;
  long long row_offset;
  int ntiles;
  float ystd, c0;  // y = ystd * f + c0, c0 = ystd * mean_const + y_mean
  // VALUES
  float* out;
  long long ld_out;
  // VALUES, row-norm variant (psbax_eval_sumsq): rowsq[chunk][N] = sum over the chunk's samples of y^2
  float* rowsq;
  int rowsq_group;  // 0: one sum per 64-sample chunk; g (a divisor of 64): rowsq[s / g][N], one sum per g consecutive samples
  // LEVELSET
  float tau;
  uint32_t* mask;
  long long mask_ld;
  // per-CTA partial results, [gridDim.x][Sp] (+ [k] for the lists)
  unsigned int* p_cnt;
  float* p_max;
  int* p_arg;
  int k;
  float* p_tv;
  int* p_ti;
  const float* thr_seed;  // TOPK, optional [Sp]: admission threshold known before the launch (see psbax_eval_topk)
  long long* trace;  // optional device buffer for clock64 traces (psbax_debug_trace_buffer)
  // LEVELSET, optional: pooled[w] |= OR over the S samples of the mask words (bit i = row i is in at
  // least one path's super-level set); zeroed by the caller, combined with atomic ORs
  uint32_t* pooled;
  // mesh mode (psbax_eval_levelset_ex): X == NULL and candidate row r is row (row_offset + r) of the
  // 'ij'-ordered mesh of the axes (first axis slowest), generated on the fly
  const float* mesh_axis[4];
  long long mesh_stride[4];
  int mesh_len[4];
  // fast path (every global row index of the call fits 32 bits): the multi-index is peeled off from the
  // last axis with multiply-high divisions by the axis lengths, q = (t + ((g - t) >> 1)) >> (shift - 1),
  // t = umulhi(g, magic), magic = floor(2^(32 + shift) / len) + 1 - 2^32, shift = ceil(log2 len)
  unsigned int mesh_magic[4];
  int mesh_shift[4];
  int mesh_fast;
};

__device__ __forceinline__ unsigned int mesh_div(const KArgs& a, unsigned int g, int ax) {
  if (a.mesh_len[ax] == 1) return g;
  const unsigned int t = __umulhi(g, a.mesh_magic[ax]);
  return (t + ((g - t) >> 1)) >> (a.mesh_shift[ax] - 1);
}

// coordinate i of local candidate row `row`: from X, or generated from the mesh axes
__device__ __forceinline__ float load_x(const KArgs& a, long long row, int i) {
  if (a.X != nullptr) return __ldg(a.X + row * a.lay.d + i);
  if (a.mesh_fast) {
    unsigned int g = (unsigned int)(a.row_offset + row);
#pragma unroll
    for (int ax = 3; ax > 0; --ax) {
      if (ax < a.lay.d && ax >= i) {
        const unsigned int q = mesh_div(a, g, ax);
        if (ax == i) return __ldg(a.mesh_axis[i] + (g - q * (unsigned int)a.mesh_len[ax]));
        g = q;
      }
    }
    return __ldg(a.mesh_axis[0] + g);  // i == 0: the slowest axis is what is left (row < prod(len))
  }
  const long long g = a.row_offset + row;
  const int j = (int)((g / a.mesh_stride[i]) % a.mesh_len[i]);
  return __ldg(a.mesh_axis[i] + j);
}

// ---------------------------------------------------------------------------------
// math helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Unit-variance stationary kernel profile of the squared scaled distance.
// kernel id is warp-uniform.  sqrt/ex2 go to the MUFU pipe.
__device__ __forceinline__ float kernel_profile(float r2, int kernel) {
  constexpr float LOG2E = 1.4426950408889634f;
  if (kernel == PSBAX_KERNEL_RBF) return fast_ex2(-0.5f * LOG2E * r2);
  float r = fast_sqrt(r2);
  if (kernel == PSBAX_KERNEL_MATERN52) {
    float s = 2.2360679774997896f * r;
    return (1.0f + s + 1.6666666666666667f * r2) * fast_ex2(-LOG2E * s);
  }
  if (kernel == PSBAX_KERNEL_MATERN32) {
    float s = 1.7320508075688772f * r;
    return (1.0f + s) * fast_ex2(-LOG2E * s);
  }
  return fast_ex2(-LOG2E * r);
}

// NP profiles at once: one (warp-uniform) switch, straight-line bodies
template <int NP>
__device__ __forceinline__ void kernel_profile_n(const float (&r2)[NP], float (&v)[NP], int kernel) {
  constexpr float LOG2E = 1.4426950408889634f;
  if (kernel == PSBAX_KERNEL_MATERN52) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float s = 2.2360679774997896f * fast_sqrt(r2[j]);
      v[j] = (1.0f + s + 1.6666666666666667f * r2[j]) * fast_ex2(-LOG2E * s);
    }
  } else if (kernel == PSBAX_KERNEL_RBF) {
#pragma unroll
    for (int j = 0; j < NP; ++j) v[j] = fast_ex2(-0.5f * LOG2E * r2[j]);
  } else if (kernel == PSBAX_KERNEL_MATERN32) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float s = 1.7320508075688772f * fast_sqrt(r2[j]);
      v[j] = (1.0f + s) * fast_ex2(-LOG2E * s);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NP; ++j) v[j] = fast_ex2(-LOG2E * fast_sqrt(r2[j]));
  }
}
__device__ __forceinline__ void kernel_profile8(const float (&r2)[8], float (&v)[8], int kernel) {
  kernel_profile_n<8>(r2, v, kernel);
}

// d kappa / d(r2) * 2  (so that grad_x = dk2 * diff * inv_ls), used by the gradient kernel
__device__ __forceinline__ float kernel_profile_dr2x2(float r2, int kernel) {
  if (kernel == PSBAX_KERNEL_RBF) return -__expf(-0.5f * r2);
  float r = sqrtf(r2);
  if (kernel == PSBAX_KERNEL_MATERN52) {
    float s = 2.2360679774997896f * r;
    return -(5.0f / 3.0f) * (1.0f + s) * __expf(-s);
  }
  if (kernel == PSBAX_KERNEL_MATERN32) return -3.0f * __expf(-1.7320508075688772f * r);
  return r > 0.f ? -__expf(-r) / r : 0.f;
}

// total order used everywhere a "best" entry is chosen: larger value first, ties
// towards the smaller index
template <typename I>
__device__ __forceinline__ bool better(float v, I i, float v2, I i2) {
  return v > v2 || (v == v2 && i < i2);
}

}  // namespace psbax
