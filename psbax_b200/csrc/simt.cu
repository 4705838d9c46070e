// Translation unit: the FFMA / MUFU kernels (k_shared_simt, k_persample) and their dispatch.
#include <cstdio>

#include "sp_simt.cuh"

namespace psbax {
namespace {

size_t simt_smem_bytes(const Layout& l, int epi, int k, bool shared, bool generic) {
  size_t state = round_up((int)epi_smem_floats(epi, k), 4);
  size_t tile = (size_t)TN * OUT_LD;
  if (shared) {
    size_t work = (size_t)TK * TM + (size_t)TK * TN + (size_t)TK * l.OS + round_up(l.dp4, 4) +
                  (generic ? (size_t)l.dp4 * TM : 0);
    if (work > tile) tile = work;
  } else {
    tile += round_up(l.dp4, 4) + (generic ? (size_t)l.dp4 * TM : 0);
  }
  return (state + tile) * sizeof(float);
}

template <typename K>
int launch_simt(K kern, const KArgs& a, dim3 grid, size_t smem, cudaStream_t st, char* err, size_t errlen) {
  if (smem > 227 * 1024) {
    snprintf(err, errlen, "shared memory %zu B exceeds 227 KB (d too large)", smem);
    return PSBAX_E_UNSUPPORTED;
  }
  cudaError_t ce = cudaSuccess;
  if (smem > 48 * 1024) ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce == cudaSuccess) {
    kern<<<grid, NTHREADS, smem, st>>>(a);
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "FFMA engine launch failed: %s", cudaGetErrorString(ce));
    return PSBAX_E_CUDA;
  }
  return PSBAX_OK;
}

template <int EPI>
int dispatch_simt(const KArgs& a, dim3 grid, cudaStream_t st, char* err, size_t errlen) {
  const Layout& l = a.lay;
  const bool shared = l.mode != MODE_PERSAMPLE;
#define SIMT_LAUNCH(KERN, SH, GENX) return launch_simt(KERN, a, grid, simt_smem_bytes(l, EPI, a.k, SH, GENX), st, err, errlen)
  if (shared) {
    switch (l.dp4) {
      case 2:  SIMT_LAUNCH((k_shared_simt<2, EPI>), true, false);
      case 4:  SIMT_LAUNCH((k_shared_simt<4, EPI>), true, false);
      case 8:  SIMT_LAUNCH((k_shared_simt<8, EPI>), true, false);
      case 12: SIMT_LAUNCH((k_shared_simt<12, EPI>), true, false);
      case 16: SIMT_LAUNCH((k_shared_simt<16, EPI>), true, false);
      default: SIMT_LAUNCH((k_shared_simt<0, EPI>), true, true);
    }
  }
  switch (l.dp4) {
    case 2:  SIMT_LAUNCH((k_persample<2, EPI>), false, false);
    case 4:  SIMT_LAUNCH((k_persample<4, EPI>), false, false);
    case 8:  SIMT_LAUNCH((k_persample<8, EPI>), false, false);
    default: SIMT_LAUNCH((k_persample<0, EPI>), false, true);
  }
#undef SIMT_LAUNCH
}

}  // namespace

int simt_launch(int epi, const KArgs& a, dim3 grid, cudaStream_t st, char* err, size_t errlen) {
  if (epi == EPI_VALUES) return dispatch_simt<EPI_VALUES>(a, grid, st, err, errlen);
  if (epi == EPI_LEVELSET) return dispatch_simt<EPI_LEVELSET>(a, grid, st, err, errlen);
  return dispatch_simt<EPI_TOPK>(a, grid, st, err, errlen);
}
}  // namespace psbax
