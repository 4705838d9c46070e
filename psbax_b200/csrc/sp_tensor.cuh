// tcgen05 (3xTF32) engine for the shared-basis path.  Placeholder until the kernel lands:
// reports "unsupported" so that PSBAX_ENGINE_AUTO resolves to the SIMT kernels.
#pragma once
#include "sp_common.cuh"

namespace psbax {

constexpr int TC_MAX_D = 16;
constexpr int TC_MAX_S = 256;

inline bool tensor_supported(const psbax_paths_t*) { return false; }
inline size_t tensor_pack_floats(const Layout&) { return 0; }
inline int tensor_parts(const Layout&, long long, int) { return 1; }
inline int tensor_pack(const Layout&, float*, cudaStream_t) { return PSBAX_E_UNSUPPORTED; }
inline int tensor_launch(int, const KArgs&, int, cudaStream_t, char* err, size_t errlen) {
  snprintf(err, errlen, "tensor engine not built");
  return PSBAX_E_UNSUPPORTED;
}

}  // namespace psbax
