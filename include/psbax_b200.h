/* psbax_b200.h — C ABI of the B200-native PS-BAX sample-path library.
 *
 * One shared object (psbax_b200/libpsbax_b200.so, sm_100a) that evaluates S GP
 * posterior sample paths over a candidate set X* [N, d] and runs the BAX
 * algorithm reduction on every path in the same pass:
 *
 *   f_s(x) = y_mean + y_std * ( mean_const
 *                              + sum_l W[s,l] * cos(omega[g,l,:] . x + bias[g,l])
 *                              + sum_j V[s,j] * kappa(|| x*inv_ls - Xtrain[j] ||) )
 *
 * g = 0 when G == 1 (shared basis, Matheron/north-star form), g = s when G == S
 * (one basis per sample: the reference's botorch `get_gp_samples` form).
 *
 * The reference has no FFI: the interface replaced is the Python callable
 * protocol of /root/reference.  Each entry point cites the
 * reference code whose work it takes over.
 *
 * Conventions
 *  - every function returns 0 on success, a negative PSBAX_E_* code otherwise and
 *    never throws; psbax_last_error() gives the message of the calling thread's
 *    last failure;
 *  - the caller owns every buffer; all pointers in psbax_paths_t and all data
 *    arguments are DEVICE pointers (fp32, row-major, contiguous) unless stated;
 *    nothing is allocated after psbax_pack_bytes()/psbax_workspace_bytes();
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    calls are asynchronous with respect to the host;
 *  - re-entrant: no global mutable state besides per-device function attributes and the calling
 *    thread's last error message.
 */
#ifndef PSBAX_B200_H
#define PSBAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSBAX_ABI_VERSION 2

enum {
  PSBAX_OK = 0,
  PSBAX_E_INVALID = -1,     /* bad argument / inconsistent shapes */
  PSBAX_E_UNSUPPORTED = -2, /* shape outside what the kernels are built for */
  PSBAX_E_CUDA = -3,        /* a CUDA runtime call failed */
  PSBAX_E_WORKSPACE = -4    /* pack / workspace buffer too small */
};

enum { /* psbax_paths_t.kernel */
  PSBAX_KERNEL_RBF = 0,
  PSBAX_KERNEL_MATERN12 = 1,
  PSBAX_KERNEL_MATERN32 = 2,
  PSBAX_KERNEL_MATERN52 = 3,
  PSBAX_KERNEL_LINEAR = 4 /* kappa(x, x') = (x*inv_ls) . Xtrain_j: the DKGP / LinearKernel path
                             (src/utils.py:153-206); no RFF basis exists for it, use L = 0 */
};

enum { /* psbax_paths_t.engine: which kernel family evaluates a shared basis.  AUTO picks
        * the tensor engines when the basis is shared (G == 1) and S >= 2 or d >= 5, else the FFMA / MUFU kernels. */
  PSBAX_ENGINE_AUTO = 0,
  PSBAX_ENGINE_SIMT = 1,   /* FFMA + MUFU kernels */
  PSBAX_ENGINE_TENSOR = 2,      /* tcgen05 kernel, 3xTF32 operand split (G == 1 only) */
  PSBAX_ENGINE_TENSOR_BF16 = 3, /* tcgen05 kernel, BF16x3 operand split (G == 1 only) */
  PSBAX_ENGINE_TENSOR_PROJ = 4  /* BF16x3 kernel whose feature arguments omega.x + b are 3xTF32 tcgen05 MMAs
                                   too (G == 1, 5 <= d <= 32); AUTO picks it whenever it applies */
};

/* The S sample paths.  Replaces the closure `get_gp_samples` returns
 * (src/utils.py:222-229; botorch 0.10 gp_sampling.get_deterministic_model). */
typedef struct psbax_paths {
  int32_t d;       /* input dimension                                         */
  int32_t L;       /* RFF features per basis (reference: 1000, utils.py:227)   */
  int32_t G;       /* number of bases: 1 (shared) or S (one per sample)        */
  int32_t S;       /* number of sample paths                                   */
  int32_t n;       /* training points in the exact-kernel update (0 = none)    */
  int32_t kernel;  /* PSBAX_KERNEL_*                                           */
  int32_t engine;  /* PSBAX_ENGINE_*                                           */
  int32_t reserved;
  const float* omega;   /* [G, L, d]  frequencies / lengthscale                */
  const float* bias;    /* [G, L]     phases in [0, 2pi)                       */
  const float* W;       /* [S, L]     weights * sqrt(2*outputscale/L)          */
  const float* Xtrain;  /* [n, d]     training inputs / lengthscale            */
  const float* inv_ls;  /* [d]        1 / lengthscale                          */
  const float* V;       /* [S, n]     kernel-update weights * outputscale      */
  float mean_const;     /* GP constant mean (standardized units)               */
  float y_mean;         /* Standardize.means  (fit_model.py:43)                */
  float y_std;          /* Standardize.stdvs                                   */
  float reserved_f;
} psbax_paths_t;

int psbax_version(void);

/* Copies the calling thread's last error message (NUL-terminated) into buf. */
int psbax_last_error(char* buf, size_t len);

/* Profiling aid, compiled in only in the trace build of the library (python -m psbax_b200.build
 * --trace, -DPSBAX_TC_TRACE=1): CTA 0 of the tensor-engine kernel records clock64() time stamps of
 * its pipeline stages in a device buffer of at least 4 * 128 * 8 int64.  The product build keeps no
 * such state and returns PSBAX_E_UNSUPPORTED. */
int psbax_debug_trace_buffer(void* device_buffer);

/* Number of SMs of the current device (sizes persistent grids / workspaces). */
int psbax_device_sms(int* sms);

/* Size of, and construction of, the device-side operand pack for one batch of
 * paths: feature tables padded/interleaved for coalesced staging, [W;V] stacked
 * K-major, and (tensor engine) the hi/lo TF32 split of [W;V] in UMMA layout.
 * Build once per draw, reuse for every evaluation. */
int psbax_pack_bytes(const psbax_paths_t* paths, size_t* bytes);
int psbax_pack(const psbax_paths_t* paths, void* pack, size_t pack_bytes, void* stream);

/* Scratch needed by the fused reductions (per-CTA partial results). */
int psbax_workspace_bytes(const psbax_paths_t* paths, int64_t N, int32_t k, size_t* bytes);

/* out[i, s] = f_s(X[i, :]), out is [N, ld_out] with ld_out >= S.
 * Replaces PosteriorMean(sample).forward / eval_gp_sample (a4) and the
 * chunk-of-100 loop eval_in_batch (src/algorithms/topk.py:48-49). */
int psbax_eval_values(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                      float* out, int64_t ld_out, void* stream);

/* out[i] = sum_s f_s(X[i, :])^2 (output units): the row-wise squared norm of the path values,
 * without the [N, S] matrix ever reaching memory.  With the pseudo-paths V = alpha * L^-1
 * (K + sigma^2 I = L L^T, no RFF part) this is the explained variance k(x,X)(K + sigma^2 I)^-1 k(X,x)
 * of the exact GP, i.e. the streaming form of the posterior variance that
 * EntropyAcquisitionFunction.forward needs for q = 1 (src/acquisition_functions/entropy.py:33-41;
 * SURVEY.md section 8, row f1).  `scratch` (psbax_sumsq_scratch_bytes; 0 bytes when S <= 64) holds
 * the per-chunk partial sums, added in ascending chunk order (deterministic). */
int psbax_sumsq_scratch_bytes(const psbax_paths_t* paths, int64_t N, size_t* bytes);
int psbax_eval_sumsq(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                     float* out, void* scratch, size_t scratch_bytes, void* stream);

/* Grouped form of psbax_eval_sumsq: out[g, i] = sum of f_s(X[i, :])^2 over the samples
 * s in [g * group, (g + 1) * group), out is [ceil(S / group), N]; `group` divides 64.  One launch then
 * yields the explained variance under SEVERAL models that share one candidate stream: the INFO-BAX
 * expected information gain (src/acquisition_functions/bax_acquisition.py:87-122, 156-174) needs the
 * posterior variance of the current model and of one fantasy model per execution path; by the block
 * Cholesky identity the fantasy of path s only adds the rows [-L_s^-1 B_s L^-1, L_s^-1] to L^-1, so all
 * of them are pseudo-paths over the union of the training set and the execution paths, one group
 * (or a few) per model. */
int psbax_eval_sumsq_groups(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                            int32_t group, float* out, void* stream);

/* Fused evaluation + LevelSetEstimator reduction (src/algorithms/levelset.py:34-39)
 * for all S paths: mask bit (i % 32) of mask[s, i / 32] = (f_s(X[i]) > tau);
 * count[s] = popcount; maxval/argmax[s] = max and (smallest) arg max of f_s over
 * the N rows, argmax reported as row_offset + i (the empty-set fallback
 * `[argmax fx]`).  mask may be NULL (argmax / ES-population scoring only);
 * mask_ld >= ceil(N/32) words per sample. */
int psbax_eval_levelset(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                        int64_t row_offset, float tau, uint32_t* mask, int64_t mask_ld,
                        int64_t* count, float* maxval, int64_t* argmax, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Candidate rows generated on the device instead of read from memory: row r of the call is row
 * (row_offset + r) of the 'ij'-ordered mesh of `d` axes, first axis slowest — the layout of
 * reshape_mesh(get_mesh(d, steps)) (src/utils.py:264-286) that demos/level-set.py:40-41 feeds the
 * level-set algorithm as x_set.  The axis values are read as given (no linspace is recomputed), so
 * the rows are bit-identical to the float32 mesh the caller would otherwise upload. */
typedef struct psbax_mesh {
  int32_t d;            /* number of axes; must equal paths->d, 1 <= d <= 4 */
  int32_t reserved;
  int64_t len[4];       /* points per axis */
  const float* axis[4]; /* DEVICE pointers: axis[i] holds len[i] floats */
} psbax_mesh_t;

/* psbax_eval_levelset with two options (either may be NULL):
 *  - mesh: the candidate rows come from a mesh descriptor (then X must be NULL and
 *    row_offset + N <= prod(len));
 *  - pooled: [ceil(N/32)] words, bit (i % 32) of pooled[i / 32] = OR over the S paths of the mask bits,
 *    i.e. candidate i is in at least one path's super-level set: the pool
 *    gen_posterior_sampling_batch builds from the S outputs
 *    (src/acquisition_functions/posterior_sampling.py:64-78) without its duplicates.  The call
 *    zeroes the buffer (asynchronously, on `stream`) before the kernel ORs into it.
 * mask may be NULL as well: only N/8 bytes then leave the kernel. */
int psbax_eval_levelset_ex(const psbax_paths_t* paths, const void* pack, const float* X,
                           const psbax_mesh_t* mesh, int64_t N, int64_t row_offset, float tau,
                           uint32_t* mask, int64_t mask_ld, uint32_t* pooled, int64_t* count,
                           float* maxval, int64_t* argmax, void* workspace, size_t workspace_bytes,
                           void* stream);

/* Fused evaluation + TopK reduction (src/algorithms/topk.py:31-34) for all S
 * paths: out_val[s, 0..k) the k largest f_s values, largest first, ties broken
 * towards the smaller index; out_idx[s, j] = row_offset + row.  1 <= k <= 64,
 * k <= N. */
int psbax_eval_topk(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                    int64_t row_offset, int32_t k, float* out_val, int64_t* out_idx,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Merge `lists` sorted top-k lists per sample (in_val/in_idx are
 * [lists, S, k]) into one [S, k] list; the step after the all-gather of the
 * per-GPU shard results.  Unused slots carry -inf / -1. */
int psbax_topk_merge(const float* in_val, const int64_t* in_idx, int32_t lists, int32_t S,
                     int32_t k, float* out_val, int64_t* out_idx, void* stream);

/* DiscoBAX SubsetSelect reduction (src/algorithms/discobax.py:37-56) for S
 * paths at once: values[b, i] = fx[i, s] + etas[b, i] (optionally max(0, .)),
 * first pick = arg max of the column means, then k-1 greedy picks maximising
 * mean_b max(cur_max[b], values[b, j]); already-chosen j score 0; ties to the
 * smaller index.  fx is [N, ld_fx] as written by psbax_eval_values, etas is
 * [B, N], out_idx is [S, k] (int64). */
int psbax_subset_select(const float* fx, int64_t ld_fx, const float* etas, int64_t N, int32_t B,
                        int32_t S, int32_t k, int32_t nonneg, int64_t* out_idx, void* stream);

/* psbax_subset_select for the objective's other noise model and for per-sample noise
 * (src/problems.py:146-163, get_noisy_f_lst): noise_mode PSBAX_NOISE_MULTIPLICATIVE scores
 * values[b, i] = fx[i, s] * etas[b, i] with etas the 0 / 1 Bernoulli masks the objective draws;
 * eta_stride_s != 0 gives every sample its own [B, N] noise matrix at etas + s * eta_stride_s (the
 * reference draws a fresh mask per get_noisy_f_lst call, i.e. per sample path), 0 shares one. */
enum { PSBAX_NOISE_ADDITIVE = 0, PSBAX_NOISE_MULTIPLICATIVE = 1 };
int psbax_subset_select_ex(const float* fx, int64_t ld_fx, const float* etas, int64_t eta_stride_s, int64_t N,
                           int32_t B, int32_t S, int32_t k, int32_t nonneg, int32_t noise_mode, int64_t* out_idx,
                           void* stream);

/* grad[i, :] = d f_{s_i}(X[i, :]) / dx and val[i] = f_{s_i}(X[i, :]) for a list of
 * (point, sample) pairs: what LBFGSB obtains through autograd
 * (src/algorithms/lbfgsb.py:29-39).  sample_idx is int32 [N] (NULL = sample 0). */
int psbax_eval_grad(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                    const int32_t* sample_idx, float* val, float* grad, void* stream);

/* The feed-forward embedding of the deep-kernel GP (FFNN, src/models/deep_kernel_gp.py:18-84; the input
 * transform of the DKGP / LinearKernel sample paths, src/utils.py:186-206): out[i, :] = net(X[i, :]) in
 * fp32, Linear + activation layers with no activation after the last one.  W[l] is [width[l], width[l+1]]
 * (the TRANSPOSE of torch.nn.Linear.weight, row-major), b[l] is [width[l+1]]; all DEVICE pointers.
 * 1 <= n_layers <= PSBAX_MLP_MAX_LAYERS, every width <= 256. */
#define PSBAX_MLP_MAX_LAYERS 8
enum { PSBAX_ACT_RELU = 0, PSBAX_ACT_LEAKY_RELU = 1, PSBAX_ACT_SWISH = 2, PSBAX_ACT_SIGMOID = 3, PSBAX_ACT_TANH = 4 };
typedef struct psbax_mlp {
  int32_t n_layers;
  int32_t activation;     /* PSBAX_ACT_* */
  int32_t width[PSBAX_MLP_MAX_LAYERS + 1];
  int32_t reserved;
  const float* W[PSBAX_MLP_MAX_LAYERS];
  const float* b[PSBAX_MLP_MAX_LAYERS];
  float negative_slope;   /* leaky_relu (torch default 0.01) */
  float reserved_f;
} psbax_mlp_t;
int psbax_mlp_forward(const psbax_mlp_t* mlp, const float* X, int64_t N, float* out, void* stream);

/* Batched multi-start ascent: P independent box-constrained ascents on [0, 1]^d, ascent i maximising
 * sample path sample_idx[i] from the start X0[i, :], all inside ONE launch (one CTA per ascent; projected
 * L-BFGS with 8 curvature pairs and Armijo backtracking, value + gradient as psbax_eval_grad).  Replaces the
 * serial scipy L-BFGS-B loop under botorch optimize_acqf that LBFGSB.run_algorithm_on_f drives
 * (src/algorithms/lbfgsb.py:24-40, src/utils.py:120-150: num_restarts = 5 d ascents per sample path,
 * maxiter 100): restarts x samples ascents per call.  Stops when the projected gradient's max norm is
 * <= pgtol, the relative increase of f is <= ftol, the line search fails, or after maxiter iterations.
 * x_out [P, d], f_out [P] (f at x_out, output units), iters [P] (may be NULL); d <= 32. */
int psbax_ascend(const psbax_paths_t* paths, const void* pack, const float* X0, int64_t P,
                 const int32_t* sample_idx, int32_t maxiter, float pgtol, float ftol, float* x_out,
                 float* f_out, int32_t* iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSBAX_B200_H */
