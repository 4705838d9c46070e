// Where should the producers' feature table live?  (profiling aid, round 2)
// The bench kernel's broadcast LDS.128 of one (omega_1, omega_2, b, pad) row per feature costs two shared-memory
// wavefronts and a 512-byte register write-back each: at 16 producer warps that is ~1000 cycles of the LSU data path
// per 64-feature block, as much as the XU pipe's own floor.  This probe times the producers' projection + cos mix
// (with and without the BF16 split) with the table read four ways:
//   0: LDS.128 per feature (the kernel today)         1: 12-byte rows, three LDS.128 per four features
//   2: kernel-parameter (constant bank) table, vector-indexed LDC      3: the same with a warp-uniform index (LDCU)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/table_probe.bin scripts/table_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NF = 1024;
struct Args { float tab[3 * NF]; float* out; int iters; };

__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t cvt2(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void resid2(uint32_t p, float v_lo, float v_hi, float& r_lo, float& r_hi) {
  asm("{\n\t.reg .b16 l, h, m;\n\tmov.b32 {l, h}, %2;\n\tmov.b16 m, 0xBF80;\n\t"
      "fma.rn.f32.bf16 %0, l, m, %3;\n\tfma.rn.f32.bf16 %1, h, m, %4;\n\t}"
      : "=f"(r_lo), "=f"(r_hi) : "r"(p), "f"(v_lo), "f"(v_hi));
}

template <int MODE, bool SPLIT, int H>
__device__ __forceinline__ void body(const Args& a, const float4* tab4, const float* tab3, uint64_t x0, uint64_t x1, uint32_t& acc) {
  for (int kb = 0; kb < a.iters; ++kb) {
    const int f0 = (kb & 15) * 64 + H * 16;
    float v[2][16];
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += 4) {
      float w[4][3];
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float4 t = tab4[f0 + j0 + j]; w[j][0] = t.x; w[j][1] = t.y; w[j][2] = t.z; }
      } else if (MODE == 1) {
        const float4* t4 = reinterpret_cast<const float4*>(tab3 + (f0 + j0) * 3);
        const float4 t0 = t4[0], t1 = t4[1], t2 = t4[2];
        const float f[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { w[j][0] = f[3 * j]; w[j][1] = f[3 * j + 1]; w[j][2] = f[3 * j + 2]; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float* t = a.tab + (f0 + j0 + j) * 3; w[j][0] = t[0]; w[j][1] = t[1]; w[j][2] = t[2]; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint64_t r = fma2(x0, pk(w[j][0], w[j][0]), pk(w[j][2], w[j][2]));
        r = fma2(x1, pk(w[j][1], w[j][1]), r);
        float a0, a1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
        v[0][j0 + j] = __cosf(a0);
        v[1][j0 + j] = __cosf(a1);
      }
    }
    if (SPLIT) {
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t p = cvt2(v[t][2 * c + 1], v[t][2 * c]);
          float r0, r1;
          resid2(p, v[t][2 * c], v[t][2 * c + 1], r0, r1);
          acc += p + cvt2(r1, r0);  // IADD3: two inputs per instruction, short chain
        }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) acc += __float_as_uint(v[0][c]) + __float_as_uint(v[1][c]);
    }
  }
}

template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(1024) kprobe(const __grid_constant__ Args a) {
  __shared__ float4 tab4[NF];
  __shared__ __align__(16) float tab3[3 * NF];
  for (int i = threadIdx.x; i < NF; i += blockDim.x) {
    tab4[i] = make_float4(a.tab[3 * i], a.tab[3 * i + 1], a.tab[3 * i + 2], 0.f);
    tab3[3 * i] = a.tab[3 * i]; tab3[3 * i + 1] = a.tab[3 * i + 1]; tab3[3 * i + 2] = a.tab[3 * i + 2];
  }
  __syncthreads();
  const uint64_t x0 = pk(threadIdx.x * 1e-3f, 0.5f), x1 = pk(2.f, 3.f + threadIdx.x);
  uint32_t acc = 0;
  const int h = (threadIdx.x >> 7) & 3;
  if (MODE == 3) {  // warp-uniform index: one code copy per feature slice, offsets are immediates
    switch (h) {
      case 0: body<MODE, SPLIT, 0>(a, tab4, tab3, x0, x1, acc); break;
      case 1: body<MODE, SPLIT, 1>(a, tab4, tab3, x0, x1, acc); break;
      case 2: body<MODE, SPLIT, 2>(a, tab4, tab3, x0, x1, acc); break;
      default: body<MODE, SPLIT, 3>(a, tab4, tab3, x0, x1, acc); break;
    }
  } else {
    // runtime slice: same code as the kernel (h enters the address)
    const float4* t4 = tab4 + h * 16;
    const float* t3 = tab3 + h * 48;
    Args const& ar = a;
    if (MODE == 2) {
      // vector-indexed constant loads
      for (int kb = 0; kb < a.iters; ++kb) {
        const int f0 = (kb & 15) * 64 + h * 16;
        float v[2][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float* t = ar.tab + (f0 + j) * 3;
          uint64_t r = fma2(x0, pk(t[0], t[0]), pk(t[2], t[2]));
          r = fma2(x1, pk(t[1], t[1]), r);
          float a0, a1;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(r));
          v[0][j] = __cosf(a0);
          v[1][j] = __cosf(a1);
        }
        if (SPLIT) {
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t p = cvt2(v[t][2 * c + 1], v[t][2 * c]);
              float r0, r1;
              resid2(p, v[t][2 * c], v[t][2 * c + 1], r0, r1);
              acc += p + cvt2(r1, r0);
            }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) acc += __float_as_uint(v[0][c]) + __float_as_uint(v[1][c]);
        }
      }
    } else {
      body<MODE, SPLIT, 0>(a, t4, t3, x0, x1, acc);
    }
  }
  a.out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

template <typename K>
void time_kernel(const char* name, K kern, int threads, int iters) {
  static Args a;
  for (int i = 0; i < 3 * NF; ++i) a.tab[i] = 0.37f * (i % 97) - 11.f;
  cudaMalloc(&a.out, 148 * 1024 * sizeof(float));
  a.iters = iters;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<148, threads>>>(a);
  cudaEventRecord(e0);
  kern<<<148, threads>>>(a);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const cudaError_t err = cudaGetLastError();
  const double ops = 148.0 * threads * iters * 32.0;
  printf("%-52s %4d thr/SM: %7.3f ms  %6.2f cos/clk/SM @1965MHz %s\n", name, threads, ms, ops / (ms * 1e-3) / 148 / 1.965e9,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
  cudaFree(a.out);
}

int main() {
  for (int th : {512, 768}) {
    time_kernel("0 LDS.128 per feature, no split", kprobe<0, false>, th, 2048);
    time_kernel("1 12-byte rows (3 LDS.128 / 4 features), no split", kprobe<1, false>, th, 2048);
    time_kernel("2 constant bank, vector index (LDC), no split", kprobe<2, false>, th, 2048);
    time_kernel("3 constant bank, uniform index (LDCU), no split", kprobe<3, false>, th, 2048);
    time_kernel("0 LDS.128 per feature, FHFMA split", kprobe<0, true>, th, 2048);
    time_kernel("1 12-byte rows, FHFMA split", kprobe<1, true>, th, 2048);
    time_kernel("2 constant bank, vector index, FHFMA split", kprobe<2, true>, th, 2048);
    time_kernel("3 constant bank, uniform index, FHFMA split", kprobe<3, true>, th, 2048);
  }
  return 0;
}
