// Translation unit: 3xTF32 instantiations of the tcgen05 engine (k_shared_tensor<*, *, false>).
#include "sp_tensor.cuh"

namespace psbax {
int tensor_launch_tf32(int epi, const KArgs& a, int sms, cudaStream_t st, char* err, size_t errlen) {
  return tensor_launch_fmt<false>(epi, a, sms, st, err, errlen);
}
}  // namespace psbax
