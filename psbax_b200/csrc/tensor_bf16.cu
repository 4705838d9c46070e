// Translation unit: BF16x3 instantiations of the tcgen05 engine (k_shared_tensor<*, *, true>).
#include "sp_tensor.cuh"

namespace psbax {
int tensor_launch_bf16(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  return tensor_launch_fmt<true>(epi, a, sms, st, err, errlen);
}
}  // namespace psbax
