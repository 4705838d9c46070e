// tcgen05.mma throughput against the number of accumulators the issue stream rotates over (profiling aid, round 2).
// One CTA per SM, one issuing warp; A from tensor memory, B from shared memory (no-swizzle K-major images of zeros),
// kind::f16 (bf16, K = 16) and kind::tf32 (K = 8), M = 128, N = 64 / 128.  Reports cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I psbax_b200/csrc -o scripts/mma_probe.bin scripts/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sp_tensor.cuh"

using namespace psbax;

template <bool BF, int N>
__global__ void __launch_bounds__(128) kmma(long long* out, int n_mma, int nacc, int adist) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 0) {
    constexpr uint32_t IDESC = BF ? umma_idesc_bf16(N) : umma_idesc_tf32(N);
    const uint32_t sbo = BF ? 64 * 2 * 8 : 32 * 4 * 8;
    const uint64_t b = umma_desc_kmajor(smem_u32(smem), 128, sbo);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        int acc = 0;
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t d = tmem + acc * N;            // accumulators at columns 0, N, 2N, ...
          const uint32_t a = tmem + 448 + (i & 7) * 8 * (adist ? 1 : 0);  // A operand columns (8 per k-step)
          if (BF) tc_mma_ts_bf16(d, a, b, IDESC, 1u);
          else tc_mma_ts(d, a, b, IDESC, 1u);
          if (++acc == nacc) acc = 0;
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool BF, int N>
void run(const char* name, int nacc) {
  long long* out;
  cudaMalloc(&out, 8);
  const int n = 4096;
  cudaFuncSetAttribute(kmma<BF, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  kmma<BF, N><<<148, 128, 32768>>>(out, n, nacc, 1);
  long long h = 0;
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  const cudaError_t e = cudaGetLastError();
  printf("%-28s N = %3d, %d accumulator(s): %6.1f cycles / MMA %s\n", name, N, nacc, (double)h / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int nacc : {1, 2, 3, 4, 6}) run<true, 64>("kind::f16 bf16 K=16", nacc);
  for (int nacc : {1, 2, 3}) run<true, 128>("kind::f16 bf16 K=16", nacc);
  for (int nacc : {1, 2, 3, 4, 6}) run<false, 64>("kind::tf32 K=8", nacc);
  for (int nacc : {1, 2, 4, 8}) run<false, 32>("kind::tf32 K=8", nacc);
  return 0;
}
