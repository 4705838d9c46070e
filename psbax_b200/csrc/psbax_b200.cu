// C-ABI entry points of libpsbax_b200.so (declared in include/psbax_b200.h).
// Host side: argument validation, operand-pack layout, launch configuration.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#define PSBAX_TU_API 1
#include "sp_common.cuh"
#include "sp_extra.cuh"
#include "sp_mlp.cuh"
#include "sp_host_kernels.cuh"
#include "sp_simt.cuh"
#include "sp_tensor.cuh"
#include "sp_tensor_proj.cuh"

namespace psbax {
// simt.cu: the FFMA / MUFU kernels
int simt_launch(int epi, const KArgs& a, dim3 grid, cudaStream_t st, char* err, size_t errlen);
}

using namespace psbax;

namespace {

thread_local char g_err[512] = "";
#if defined(PSBAX_TC_TRACE) && PSBAX_TC_TRACE
long long* g_trace = nullptr;  // trace build only, see psbax_debug_trace_buffer
#endif

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return fail(PSBAX_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

int device_sms(int* sms) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return PSBAX_OK;
}

int validate(const psbax_paths_t* p) {
  if (!p) return fail(PSBAX_E_INVALID, "paths is NULL");
  if (p->d < 1 || p->d > 1024) return fail(PSBAX_E_INVALID, "d=%d out of range [1,1024]", p->d);
  if (p->S < 1) return fail(PSBAX_E_INVALID, "S=%d must be >= 1", p->S);
  if (p->L < 0 || p->n < 0) return fail(PSBAX_E_INVALID, "L=%d, n=%d must be >= 0", p->L, p->n);
  if (p->L + p->n < 1) return fail(PSBAX_E_INVALID, "L + n must be >= 1");
  if (p->G != 1 && p->G != p->S)
    return fail(PSBAX_E_INVALID, "G=%d must be 1 (shared basis) or S=%d", p->G, p->S);
  if (p->kernel < PSBAX_KERNEL_RBF || p->kernel > PSBAX_KERNEL_LINEAR)
    return fail(PSBAX_E_INVALID, "unknown kernel id %d", p->kernel);
  if (p->L > 0 && (!p->omega || !p->bias || !p->W))
    return fail(PSBAX_E_INVALID, "omega / bias / W must be non-NULL when L > 0");
  if (p->n > 0 && (!p->Xtrain || !p->inv_ls || !p->V))
    return fail(PSBAX_E_INVALID, "Xtrain / inv_ls / V must be non-NULL when n > 0");
  if (p->engine < PSBAX_ENGINE_AUTO || p->engine > PSBAX_ENGINE_TENSOR_PROJ)
    return fail(PSBAX_E_INVALID, "unknown engine id %d", p->engine);
  if ((p->engine == PSBAX_ENGINE_TENSOR && !tensor_supported(p, false)) ||
      (p->engine == PSBAX_ENGINE_TENSOR_BF16 && !tensor_supported(p, true)))
    return fail(PSBAX_E_UNSUPPORTED,
                "tensor engine needs a shared basis (G == 1), L + n > 0, d <= %d and operands that fit "
                "shared memory (got G=%d d=%d S=%d L=%d n=%d)", TC_MAX_D, p->G, p->d, p->S, p->L, p->n);
  if (p->engine == PSBAX_ENGINE_TENSOR_PROJ && !tensor_proj_supported(p))
    return fail(PSBAX_E_UNSUPPORTED,
                "tensor projection engine needs a shared basis (G == 1), L + n > 0, %d <= d <= %d and operands "
                "that fit shared memory (got G=%d d=%d S=%d L=%d n=%d)", TP_MIN_D, TP_MAX_D, p->G, p->d, p->S, p->L, p->n);
  return PSBAX_OK;
}

Layout make_layout(const psbax_paths_t* p) {
  Layout l{};
  l.d = p->d; l.L = p->L; l.G = p->G; l.S = p->S; l.n = p->n; l.kernel = p->kernel;
  l.dp4 = (p->d <= 2) ? 2 : round_up(p->d, 4);
  l.OS = round_up(l.dp4 + 1, 4);
  l.OSB = round_up(l.dp4 + 2, 4);
  l.Lp = round_up(p->L, 8);
  l.np_ = round_up(p->n, 8);
  l.Lcos = l.Lp;
  if (p->kernel == PSBAX_KERNEL_LINEAR) {  // the n linear rows join the feature section
    l.lin_n = p->n;
    l.n = 0;
    l.np_ = 0;
    l.Lp = l.Lcos + round_up(p->n, 8);
  }
  l.Sp = round_up(p->S, TN);
  // One shared basis: generate every feature once per candidate and contract with all samples (the
  // sample dimension is padded to 64 anyway).  Only a single path over a low-dimensional set is
  // better off in the per-sample kernel (measured at N = 1e6, d = 2: S = 1 0.39 ms there against
  // 0.50 ms here; S = 4 / 8 / 15 took 1.3 / 2.7 / 5.2 ms there, d = 10 with S = 1 22 ms).
  const bool tensor_asked = p->engine == PSBAX_ENGINE_TENSOR || p->engine == PSBAX_ENGINE_TENSOR_BF16 ||
                            p->engine == PSBAX_ENGINE_TENSOR_PROJ;
  const bool shared = (p->G == 1) && (l.Lp + l.np_ > 0) && (p->S >= 2 || p->d >= 5 || tensor_asked);
  if (!shared) l.mode = MODE_PERSAMPLE;
  else if (p->engine == PSBAX_ENGINE_TENSOR) l.mode = MODE_TENSOR;
  else if (p->engine == PSBAX_ENGINE_TENSOR_BF16 || p->engine == PSBAX_ENGINE_TENSOR_PROJ ||
           (p->engine == PSBAX_ENGINE_AUTO && tensor_supported(p, true)))
    l.mode = MODE_TENSOR_BF16;  // AUTO: the faster split; both hold the 1e-4 contract
  else l.mode = MODE_SHARED_SIMT;
  l.Lp_ft = (l.mode == MODE_PERSAMPLE) ? 0 : l.Lp;
  l.tkb = TK;
  if (is_tensor_mode(l.mode)) {
    const bool want_proj = p->engine == PSBAX_ENGINE_TENSOR_PROJ ||
                           (p->engine == PSBAX_ENGINE_AUTO && p->d >= TP_AUTO_MIN_D);
    l.tkb = TP_TKB;
    layout_blocks(l);  // what the projection variant would use (tp_supported reads Kp)
    if (want_proj && tp_supported(l)) {  // [W;V] images as for the generic kernel (32-feature bf16 blocks)
      l.tc_proj = 1;
      l.tc_generic = 1;
    } else {
      l.tc_generic = tc_choose_generic(l) ? 1 : 0;
    }
    l.tkb = tc_tkb(l.mode == MODE_TENSOR_BF16, l.tc_generic != 0);
  }
  layout_blocks(l);
  size_t off = 0;
  l.off_invls = off; off += round_up(l.dp4, 4);
  l.off_ft = off;    off += (size_t)l.Kp * l.OS;
  l.off_wv = off;    off += (size_t)l.Kp * l.Sp;
  l.off_pb = off;    if (l.mode == MODE_PERSAMPLE) off += (size_t)l.S * l.Lp * l.OSB;
  off = (off + 63) / 64 * 64;  // 256-byte alignment for the tensor operands
  l.off_tc = off;    if (is_tensor_mode(l.mode)) off += tensor_pack_floats(l);
  if (l.tc_proj) off += tp_pack_floats(l);  // Omega_aug images behind the [W;V] images
  l.total_floats = off;
  return l;
}

// fused top-k: [Sp] seed thresholds + [TOPK_SEED_ROWS][Sp] pre-pass values (see psbax_eval_topk)
size_t topk_seed_bytes(const Layout& l) { return ((size_t)l.Sp + (size_t)TOPK_SEED_ROWS * l.Sp) * sizeof(float) + 256; }

// persistent grid: gx CTAs over the row tiles, gy sample chunks
void grid_for(const Layout& l, long long N, int sms, int* gx, int* gy, int* ntiles) {
  const long long nt = (N + TM - 1) / TM;
  *ntiles = (int)nt;
  *gy = l.Sp / TN;
  long long want = (2LL * sms) / *gy;  // two CTAs per SM, one wave
  if (want < 1) want = 1;
  long long g = nt < want ? nt : want;
  if (g < 1) g = 1;
  const long long tpc = (nt + g - 1) / g;
  g = (nt + tpc - 1) / (tpc > 0 ? tpc : 1);
  if (g < 1) g = 1;
  *gx = (int)g;
}

struct Prepared {
  KArgs a;
  dim3 grid;
  int sms;
};

int prepare(const psbax_paths_t* p, const void* pack, const float* X, int64_t N, Prepared* out,
            bool x_optional = false) {
  int rc = validate(p);
  if (rc) return rc;
  if (!pack) return fail(PSBAX_E_INVALID, "pack is NULL (call psbax_pack first)");
  if (!X && !x_optional) return fail(PSBAX_E_INVALID, "X is NULL");
  if (N < 1) return fail(PSBAX_E_INVALID, "N=%lld must be >= 1", (long long)N);
  if (N > 2147483647LL - TM) return fail(PSBAX_E_UNSUPPORTED, "N=%lld exceeds 2^31 rows per call; shard the candidate set", (long long)N);
  rc = device_sms(&out->sms);
  if (rc) return rc;
  KArgs& a = out->a;
  memset(&a, 0, sizeof(a));
  a.lay = make_layout(p);
  a.pack = static_cast<const float*>(pack);
  a.X = X;
  a.N = N;
  a.ystd = p->y_std;
  a.c0 = (float)((double)p->y_std * (double)p->mean_const + (double)p->y_mean);
#if defined(PSBAX_TC_TRACE) && PSBAX_TC_TRACE
  a.trace = g_trace;
#endif
  int gx, gy, nt;
  grid_for(a.lay, N, out->sms, &gx, &gy, &nt);
  a.ntiles = nt;
  out->grid = dim3(gx, gy, 1);
  return PSBAX_OK;
}

}  // namespace

extern "C" {

int psbax_version(void) { return PSBAX_ABI_VERSION; }

int psbax_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return PSBAX_E_INVALID;
  strncpy(buf, g_err, len - 1);
  buf[len - 1] = 0;
  return PSBAX_OK;
}

int psbax_debug_trace_buffer(void* device_buffer) {
#if defined(PSBAX_TC_TRACE) && PSBAX_TC_TRACE
  g_trace = static_cast<long long*>(device_buffer);
  return PSBAX_OK;
#else
  (void)device_buffer;
  return fail(PSBAX_E_UNSUPPORTED, "clock traces exist only in the trace build (python -m psbax_b200.build --trace)");
#endif
}

int psbax_device_sms(int* sms) {
  if (!sms) return fail(PSBAX_E_INVALID, "sms is NULL");
  return device_sms(sms);
}

int psbax_pack_bytes(const psbax_paths_t* paths, size_t* bytes) {
  int rc = validate(paths);
  if (rc) return rc;
  if (!bytes) return fail(PSBAX_E_INVALID, "bytes is NULL");
  *bytes = make_layout(paths).total_floats * sizeof(float);
  return PSBAX_OK;
}

int psbax_pack(const psbax_paths_t* paths, void* pack, size_t pack_bytes, void* stream) {
  int rc = validate(paths);
  if (rc) return rc;
  if (!pack) return fail(PSBAX_E_INVALID, "pack is NULL");
  const Layout l = make_layout(paths);
  if (pack_bytes < l.total_floats * sizeof(float))
    return fail(PSBAX_E_WORKSPACE, "pack buffer too small: %zu < %zu bytes", pack_bytes,
                l.total_floats * sizeof(float));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_pack<<<296, 256, 0, st>>>(l, paths->omega, paths->bias, paths->W, paths->Xtrain,
                              paths->inv_ls, paths->V, static_cast<float*>(pack));
  CUDA_TRY(cudaGetLastError());
  if (is_tensor_mode(l.mode)) {
    rc = tensor_pack(l, static_cast<float*>(pack), st);
    if (rc == PSBAX_OK && l.tc_proj) rc = tensor_proj_pack(l, static_cast<float*>(pack), st);
    if (rc) return fail(rc, "tensor operand pack failed");
  }
  return PSBAX_OK;
}

int psbax_workspace_bytes(const psbax_paths_t* paths, int64_t N, int32_t k, size_t* bytes) {
  int rc = validate(paths);
  if (rc) return rc;
  if (!bytes) return fail(PSBAX_E_INVALID, "bytes is NULL");
  if (N < 1) return fail(PSBAX_E_INVALID, "N must be >= 1");
  if (k < 0 || k > KMAX) return fail(PSBAX_E_UNSUPPORTED, "k=%d outside [0,%d]", k, KMAX);
  int sms = 0;
  rc = device_sms(&sms);
  if (rc) return rc;
  const Layout l = make_layout(paths);
  int gx, gy, nt;
  grid_for(l, N, sms, &gx, &gy, &nt);
  if (is_tensor_mode(l.mode)) gx = l.tc_proj ? sms : tensor_parts(l, N, sms);  // (upper bound for the projection variant's plans)
  const size_t per = (size_t)gx * l.Sp;
  const size_t ls = per * (sizeof(unsigned int) + sizeof(float) + sizeof(int));
  size_t tk = per * (size_t)(k > 0 ? k : 1) * (sizeof(float) + sizeof(int));
  if (k > 0) tk += topk_seed_bytes(l);  // seed thresholds + the values of the pre-pass rows
  *bytes = (ls > tk ? ls : tk) + 256;
  return PSBAX_OK;
}

int psbax_eval_values(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                      float* out, int64_t ld_out, void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X, N, &pr);
  if (rc) return rc;
  if (!out) return fail(PSBAX_E_INVALID, "out is NULL");
  if (ld_out < paths->S) return fail(PSBAX_E_INVALID, "ld_out=%lld < S=%d", (long long)ld_out, paths->S);
  pr.a.out = out;
  pr.a.ld_out = ld_out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_tensor_mode(pr.a.lay.mode)) return tensor_launch_any(EPI_VALUES, pr.a, pr.sms, st, g_err, sizeof(g_err));
  return simt_launch(EPI_VALUES, pr.a, pr.grid, st, g_err, sizeof(g_err));
}

int psbax_sumsq_scratch_bytes(const psbax_paths_t* paths, int64_t N, size_t* bytes) {
  int rc = validate(paths);
  if (rc) return rc;
  if (!bytes) return fail(PSBAX_E_INVALID, "bytes is NULL");
  if (N < 1) return fail(PSBAX_E_INVALID, "N must be >= 1");
  const int chunks = round_up(paths->S, TN) / TN;
  *bytes = chunks > 1 ? (size_t)chunks * (size_t)N * sizeof(float) : 0;
  return PSBAX_OK;
}

int psbax_eval_sumsq(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                     float* out, void* scratch, size_t scratch_bytes, void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X, N, &pr);
  if (rc) return rc;
  if (!out) return fail(PSBAX_E_INVALID, "out is NULL");
  size_t need = 0;
  rc = psbax_sumsq_scratch_bytes(paths, N, &need);
  if (rc) return rc;
  if (need > 0 && (!scratch || scratch_bytes < need))
    return fail(PSBAX_E_WORKSPACE, "scratch too small: %zu < %zu bytes", scratch_bytes, need);
  const int chunks = pr.a.lay.Sp / TN;
  pr.a.rowsq = chunks > 1 ? static_cast<float*>(scratch) : out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_tensor_mode(pr.a.lay.mode)) rc = tensor_launch_any(EPI_VALUES, pr.a, pr.sms, st, g_err, sizeof(g_err));
  else rc = simt_launch(EPI_VALUES, pr.a, pr.grid, st, g_err, sizeof(g_err));
  if (rc) return rc;
  if (chunks > 1) {
    k_sumsq_finalize<<<(unsigned)((N + 255) / 256 > 2368 ? 2368 : (N + 255) / 256), 256, 0, st>>>(
        pr.a.rowsq, chunks, N, out);
    CUDA_TRY(cudaGetLastError());
  }
  return PSBAX_OK;
}

int psbax_eval_sumsq_groups(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                            int32_t group, float* out, void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X, N, &pr);
  if (rc) return rc;
  if (!out) return fail(PSBAX_E_INVALID, "out is NULL");
  if (group < 1 || group > TN || TN % group != 0)
    return fail(PSBAX_E_INVALID, "group=%d must divide %d", group, TN);
  pr.a.rowsq = out;
  pr.a.rowsq_group = group;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_tensor_mode(pr.a.lay.mode)) return tensor_launch_any(EPI_VALUES, pr.a, pr.sms, st, g_err, sizeof(g_err));
  return simt_launch(EPI_VALUES, pr.a, pr.grid, st, g_err, sizeof(g_err));
}

int psbax_eval_levelset(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                        int64_t row_offset, float tau, uint32_t* mask, int64_t mask_ld,
                        int64_t* count, float* maxval, int64_t* argmax, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return psbax_eval_levelset_ex(paths, pack, X, nullptr, N, row_offset, tau, mask, mask_ld, nullptr, count,
                                maxval, argmax, workspace, workspace_bytes, stream);
}

int psbax_eval_levelset_ex(const psbax_paths_t* paths, const void* pack, const float* X,
                           const psbax_mesh_t* mesh, int64_t N, int64_t row_offset, float tau,
                           uint32_t* mask, int64_t mask_ld, uint32_t* pooled, int64_t* count,
                           float* maxval, int64_t* argmax, void* workspace, size_t workspace_bytes,
                           void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X, N, &pr, mesh != nullptr);
  if (rc) return rc;
  if (mask && mask_ld < (N + 31) / 32)
    return fail(PSBAX_E_INVALID, "mask_ld=%lld < ceil(N/32)=%lld", (long long)mask_ld, (long long)((N + 31) / 32));
  size_t need = 0;
  rc = psbax_workspace_bytes(paths, N, 0, &need);
  if (rc) return rc;
  if (!workspace || workspace_bytes < need)
    return fail(PSBAX_E_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  KArgs& a = pr.a;
  const Layout& l = a.lay;
  if (mesh) {
    if (X) return fail(PSBAX_E_INVALID, "pass either X or a mesh, not both");
    if (mesh->d != paths->d || mesh->d < 1 || mesh->d > 4)
      return fail(PSBAX_E_UNSUPPORTED, "mesh needs 1 <= d <= 4 axes and d == paths->d (got %d, paths->d = %d)",
                  mesh->d, paths->d);
    if (row_offset < 0) return fail(PSBAX_E_INVALID, "row_offset must be >= 0 in mesh mode");
    long long total = 1;
    for (int i = mesh->d - 1; i >= 0; --i) {
      if (mesh->len[i] < 1 || mesh->len[i] > 2147483647LL || !mesh->axis[i])
        return fail(PSBAX_E_INVALID, "mesh axis %d: len=%lld, pointer %p", i, (long long)mesh->len[i],
                    (const void*)mesh->axis[i]);
      a.mesh_axis[i] = mesh->axis[i];
      a.mesh_len[i] = (int)mesh->len[i];
      a.mesh_stride[i] = total;
      int sh = 0;
      while ((1LL << sh) < mesh->len[i]) ++sh;  // ceil(log2 len)
      a.mesh_shift[i] = sh;
      a.mesh_magic[i] = mesh->len[i] > 1 ? (unsigned int)(((1ULL << (32 + sh)) / (unsigned long long)mesh->len[i]) + 1ULL -
                                                          (1ULL << 32)) : 0u;
      if (total > (1LL << 62) / mesh->len[i]) return fail(PSBAX_E_INVALID, "mesh too large");
      total *= mesh->len[i];
    }
    if (row_offset + N > total)
      return fail(PSBAX_E_INVALID, "rows [%lld, %lld) exceed the mesh (%lld rows)", (long long)row_offset,
                  (long long)(row_offset + N), total);
    // rows past N inside the last tile are never generated (callers test row < N first)
    a.mesh_fast = (row_offset + N <= 0xFFFFFFFFLL) ? 1 : 0;
  }
  const int parts = is_tensor_mode(l.mode) ? tensor_parts_any(l, N, pr.sms, EPI_LEVELSET, 0) : (int)pr.grid.x;
  const size_t per = (size_t)parts * l.Sp;
  a.row_offset = row_offset;
  a.tau = tau;
  a.mask = mask;
  a.mask_ld = mask_ld;
  a.pooled = pooled;
  a.p_cnt = static_cast<unsigned int*>(workspace);
  a.p_max = reinterpret_cast<float*>(a.p_cnt + per);
  a.p_arg = reinterpret_cast<int*>(a.p_max + per);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pooled) CUDA_TRY(cudaMemsetAsync(pooled, 0, (size_t)((N + 31) / 32) * sizeof(uint32_t), st));
  if (is_tensor_mode(l.mode)) rc = tensor_launch_any(EPI_LEVELSET, a, pr.sms, st, g_err, sizeof(g_err));
  else rc = simt_launch(EPI_LEVELSET, a, pr.grid, st, g_err, sizeof(g_err));
  if (rc) return rc;
  k_levelset_finalize<<<(l.S + 127) / 128, 128, 0, st>>>(a.p_cnt, a.p_max, a.p_arg, parts, l.Sp, l.S,
                                                       row_offset, reinterpret_cast<long long*>(count),
                                                       maxval, reinterpret_cast<long long*>(argmax));
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_eval_topk(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                    int64_t row_offset, int32_t k, float* out_val, int64_t* out_idx,
                    void* workspace, size_t workspace_bytes, void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X, N, &pr);
  if (rc) return rc;
  if (k < 1 || k > KMAX) return fail(PSBAX_E_UNSUPPORTED, "k=%d outside [1,%d]", k, KMAX);
  if (k > N) return fail(PSBAX_E_INVALID, "k=%d > N=%lld (torch.topk raises too)", k, (long long)N);
  if (!out_val || !out_idx) return fail(PSBAX_E_INVALID, "out_val / out_idx is NULL");
  size_t need = 0;
  rc = psbax_workspace_bytes(paths, N, k, &need);
  if (rc) return rc;
  if (!workspace || workspace_bytes < need)
    return fail(PSBAX_E_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  KArgs& a = pr.a;
  const Layout& l = a.lay;
  const int parts = is_tensor_mode(l.mode) ? tensor_parts_any(l, N, pr.sms, EPI_TOPK, k) : (int)pr.grid.x;
  const size_t per = (size_t)parts * l.Sp * k;
  a.row_offset = row_offset;
  a.k = k;
  a.p_tv = static_cast<float*>(workspace);
  a.p_ti = reinterpret_cast<int*>(a.p_tv + per);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Seed thresholds.  Every CTA starts with empty lists, so its first tiles send nearly every row
  // through the list merge — with few tiles per CTA (C3: 149,361 rows = 8 tiles per SM) that cold
  // start cost more than the contraction itself.  A pre-pass evaluates the first 2048 rows (values
  // kernel, same arithmetic) and takes each sample's k-th largest value there: a lower bound of the
  // k-th largest over all N rows, so the main launch only ever merges ~N k / 2048 rows per sample.
  {
    const long long tiles_per_cta = (N + TM - 1) / TM / (parts > 0 ? parts : 1);
    if (N >= 8LL * TOPK_SEED_ROWS && tiles_per_cta < 128 && k <= TOPK_SEED_ROWS / 8) {
      float* seed = reinterpret_cast<float*>(reinterpret_cast<char*>(a.p_ti + per) + 128);
      seed = reinterpret_cast<float*>(((uintptr_t)seed + 127) & ~(uintptr_t)127);
      float* svals = seed + l.Sp;
      Prepared pv;
      rc = prepare(paths, pack, X, TOPK_SEED_ROWS, &pv);
      if (rc) return rc;
      pv.a.out = svals;
      pv.a.ld_out = l.Sp;
      if (is_tensor_mode(l.mode)) rc = tensor_launch_any(EPI_VALUES, pv.a, pv.sms, st, g_err, sizeof(g_err));
      else rc = simt_launch(EPI_VALUES, pv.a, pv.grid, st, g_err, sizeof(g_err));
      if (rc) return rc;
      k_topk_seed<<<l.S, 256, 0, st>>>(svals, TOPK_SEED_ROWS, l.Sp, k, seed);
      CUDA_TRY(cudaGetLastError());
      a.thr_seed = seed;
    }
  }
  if (is_tensor_mode(l.mode)) rc = tensor_launch_any(EPI_TOPK, a, pr.sms, st, g_err, sizeof(g_err));
  else rc = simt_launch(EPI_TOPK, a, pr.grid, st, g_err, sizeof(g_err));
  if (rc) return rc;
  k_topk_merge<int><<<l.S, 128, 0, st>>>(a.p_tv, a.p_ti, parts, l.Sp, k, row_offset, out_val,
                                        reinterpret_cast<long long*>(out_idx));
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_topk_merge(const float* in_val, const int64_t* in_idx, int32_t lists, int32_t S,
                     int32_t k, float* out_val, int64_t* out_idx, void* stream) {
  if (!in_val || !in_idx || !out_val || !out_idx) return fail(PSBAX_E_INVALID, "NULL buffer");
  if (lists < 1 || S < 1 || k < 1) return fail(PSBAX_E_INVALID, "lists=%d S=%d k=%d must be >= 1", lists, S, k);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_topk_merge<long long><<<S, 128, 0, st>>>(in_val, reinterpret_cast<const long long*>(in_idx), lists,
                                              S, k, 0LL, out_val, reinterpret_cast<long long*>(out_idx));
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_subset_select(const float* fx, int64_t ld_fx, const float* etas, int64_t N, int32_t B,
                        int32_t S, int32_t k, int32_t nonneg, int64_t* out_idx, void* stream) {
  if (!fx || !etas || !out_idx) return fail(PSBAX_E_INVALID, "NULL buffer");
  if (N < 1 || B < 1 || S < 1 || k < 1) return fail(PSBAX_E_INVALID, "N, B, S, k must be >= 1");
  if (k > N) return fail(PSBAX_E_INVALID, "k=%d > N=%lld", k, (long long)N);
  if (ld_fx < S) return fail(PSBAX_E_INVALID, "ld_fx < S");
  if (N > 2147483647LL) return fail(PSBAX_E_UNSUPPORTED, "N exceeds 2^31");
  int sms = 0;
  int rc = device_sms(&sms);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // T samples per CTA share every eta load (the L2 -> SM traffic for eta is the second bound after
  // the ALU work); one CTA of 512 threads per SM, so pick the smallest T that fits S into one wave
  const int per_sm = (S + sms - 1) / sms;
  const int T = per_sm > 8 ? 8 : per_sm;  // e.g. S = 1024 on 148 SMs: 147 CTAs of 7 samples, not 128 of 8
  const int threads = (N >= 4096) ? 512 : 256;
  const size_t smem = (size_t)T * ((B + 3) & ~3) * 4 + (size_t)T * k * 4 + (size_t)T * 32 * 8;
  if (smem > 200 * 1024) return fail(PSBAX_E_UNSUPPORTED, "B=%d too large for shared memory", B);
  const int grid = (S + T - 1) / T;
#define LAUNCH_SS2(TT, NN)                                                                      \
  do {                                                                                          \
    if (smem > 48 * 1024)                                                                       \
      CUDA_TRY(cudaFuncSetAttribute(k_subset_select<TT, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_subset_select<TT, NN><<<grid, threads, smem, st>>>(fx, ld_fx, etas, N, B, S, k,            \
                                                     reinterpret_cast<long long*>(out_idx));    \
  } while (0)
#define LAUNCH_SS(TT) do { if (nonneg) LAUNCH_SS2(TT, true); else LAUNCH_SS2(TT, false); } while (0)
  switch (T) {
    case 8: LAUNCH_SS(8); break;
    case 7: LAUNCH_SS(7); break;
    case 6: LAUNCH_SS(6); break;
    case 5: LAUNCH_SS(5); break;
    case 4: LAUNCH_SS(4); break;
    case 3: LAUNCH_SS(3); break;
    case 2: LAUNCH_SS(2); break;
    default: LAUNCH_SS(1); break;
  }
#undef LAUNCH_SS2
#undef LAUNCH_SS
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_subset_select_ex(const float* fx, int64_t ld_fx, const float* etas, int64_t eta_stride_s, int64_t N,
                           int32_t B, int32_t S, int32_t k, int32_t nonneg, int32_t noise_mode, int64_t* out_idx,
                           void* stream) {
  if (noise_mode == PSBAX_NOISE_ADDITIVE && eta_stride_s == 0)
    return psbax_subset_select(fx, ld_fx, etas, N, B, S, k, nonneg, out_idx, stream);
  if (noise_mode != PSBAX_NOISE_ADDITIVE && noise_mode != PSBAX_NOISE_MULTIPLICATIVE)
    return fail(PSBAX_E_INVALID, "noise_mode=%d", noise_mode);
  if (!fx || !etas || !out_idx) return fail(PSBAX_E_INVALID, "NULL buffer");
  if (N < 1 || B < 1 || S < 1 || k < 1) return fail(PSBAX_E_INVALID, "N, B, S, k must be >= 1");
  if (k > N) return fail(PSBAX_E_INVALID, "k=%d > N=%lld", k, (long long)N);
  if (ld_fx < S) return fail(PSBAX_E_INVALID, "ld_fx < S");
  if (N > 2147483647LL) return fail(PSBAX_E_UNSUPPORTED, "N exceeds 2^31");
  if (eta_stride_s != 0 && eta_stride_s < (int64_t)B * N) return fail(PSBAX_E_INVALID, "eta_stride_s < B * N");
  const size_t smem = ((size_t)B + k) * 4;
  if (smem > 40 * 1024) return fail(PSBAX_E_UNSUPPORTED, "B=%d too large for shared memory", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* oi = reinterpret_cast<long long*>(out_idx);
  const bool mul = noise_mode == PSBAX_NOISE_MULTIPLICATIVE;
  if (mul && nonneg) k_subset_select_general<true, true><<<S, 256, smem, st>>>(fx, ld_fx, etas, eta_stride_s, N, B, k, oi);
  else if (mul) k_subset_select_general<true, false><<<S, 256, smem, st>>>(fx, ld_fx, etas, eta_stride_s, N, B, k, oi);
  else if (nonneg) k_subset_select_general<false, true><<<S, 256, smem, st>>>(fx, ld_fx, etas, eta_stride_s, N, B, k, oi);
  else k_subset_select_general<false, false><<<S, 256, smem, st>>>(fx, ld_fx, etas, eta_stride_s, N, B, k, oi);
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_eval_grad(const psbax_paths_t* paths, const void* pack, const float* X, int64_t N,
                    const int32_t* sample_idx, float* val, float* grad, void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X, N, &pr);
  if (rc) return rc;
  if (!grad) return fail(PSBAX_E_INVALID, "grad is NULL");
  const Layout& l = pr.a.lay;
  const size_t smem = ((size_t)2 * l.dp4 + l.Lp + l.np_ + 8) * sizeof(float);
  if (smem > 200 * 1024) return fail(PSBAX_E_UNSUPPORTED, "L + n too large for the gradient kernel");
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_eval_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_eval_grad<<<(unsigned)N, 128, smem, st>>>(pr.a, sample_idx, val, grad);
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_ascend(const psbax_paths_t* paths, const void* pack, const float* X0, int64_t P,
                 const int32_t* sample_idx, int32_t maxiter, float pgtol, float ftol, float* x_out,
                 float* f_out, int32_t* iters, void* stream) {
  Prepared pr;
  int rc = prepare(paths, pack, X0, P, &pr);
  if (rc) return rc;
  if (!x_out || !f_out) return fail(PSBAX_E_INVALID, "x_out / f_out is NULL");
  if (paths->d > ASC_DMAX) return fail(PSBAX_E_UNSUPPORTED, "ascent kernel needs d <= %d (got %d)", ASC_DMAX, paths->d);
  if (maxiter < 0) return fail(PSBAX_E_INVALID, "maxiter must be >= 0");
  const Layout& l = pr.a.lay;
  const size_t smem = ((size_t)2 * l.dp4 + l.Lp + l.np_ + 8) * sizeof(float);
  if (smem > 200 * 1024) return fail(PSBAX_E_UNSUPPORTED, "L + n too large for the ascent kernel");
  if (smem > 40 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(k_ascend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_ascend<<<(unsigned)P, 128, smem, st>>>(pr.a, sample_idx, x_out, f_out, iters, maxiter, pgtol, ftol);
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

int psbax_mlp_forward(const psbax_mlp_t* mlp, const float* X, int64_t N, float* out, void* stream) {
  if (!mlp || !X || !out) return fail(PSBAX_E_INVALID, "mlp / X / out is NULL");
  if (N < 1) return fail(PSBAX_E_INVALID, "N must be >= 1");
  if (mlp->n_layers < 1 || mlp->n_layers > PSBAX_MLP_MAX_LAYERS)
    return fail(PSBAX_E_INVALID, "n_layers=%d outside [1,%d]", mlp->n_layers, PSBAX_MLP_MAX_LAYERS);
  if (mlp->activation < PSBAX_ACT_RELU || mlp->activation > PSBAX_ACT_TANH)
    return fail(PSBAX_E_UNSUPPORTED, "activation id %d", mlp->activation);
  MlpArgs a{};
  a.n_layers = mlp->n_layers;
  a.activation = mlp->activation;
  a.slope = mlp->negative_slope;
  a.X = X;
  a.out = out;
  a.N = N;
  for (int l = 0; l <= mlp->n_layers; ++l) {
    if (mlp->width[l] < 1 || mlp->width[l] > MLP_MAXW)
      return fail(PSBAX_E_UNSUPPORTED, "layer width %d outside [1,%d]", mlp->width[l], MLP_MAXW);
    a.width[l] = mlp->width[l];
    if (mlp->width[l] > a.maxw) a.maxw = mlp->width[l];
    if (l < mlp->n_layers) {
      if (!mlp->W[l] || !mlp->b[l]) return fail(PSBAX_E_INVALID, "layer %d: W / b is NULL", l);
      a.W[l] = mlp->W[l];
      a.b[l] = mlp->b[l];
    }
  }
  a.maxw = round_up(a.maxw, 4);
  const size_t smem = ((size_t)2 * a.maxw * MLP_TM + MLP_WCHUNK) * sizeof(float);
  CUDA_TRY(cudaFuncSetAttribute(k_mlp_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mlp_forward<<<(unsigned)((N + MLP_TM - 1) / MLP_TM), MLP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a);
  CUDA_TRY(cudaGetLastError());
  return PSBAX_OK;
}

}  // extern "C"
