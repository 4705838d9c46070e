// Translation unit: the projection variant of the tcgen05 engine (k_tensor_proj).
#define PSBAX_TU_PROJ 1
#include "sp_tensor_proj.cuh"

namespace psbax {
int tensor_proj_launch(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  return tensor_proj_launch_impl(epi, a, sms, st, err, errlen);
}
}  // namespace psbax
