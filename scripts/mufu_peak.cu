// MUFU.COS / packed-FMA peak probe (profiling aid): what one SM sustains when nothing else competes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mufu_peak.bin scripts/mufu_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x * 1e-3f + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) a[j] = __cosf(a[j]);                        // FMUL.RZ + MUFU.COS
      if (MODE == 1) a[j] = __cosf(a[j]) * 1.0001f + 0.5f;       // + one FFMA
      if (MODE == 2) a[j] = exp2f(a[j] * 0.001f);                // FMUL + MUFU.EX2
      if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));  // MUFU.EX2 alone
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the BF16x3 producers' instruction mix for 16 evaluations (8 features x 2 rows), nothing else:
// 8 LDS.128, 16 FFMA2 (projection), 16 FMUL.RZ + 16 MUFU.COS, 16 F2FP, 8 IMAD.U32 + 8 LOP3, 8 FFMA2
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned cvt2(float a, float b) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__global__ void __launch_bounds__(1024) kmix(float* out, int iters, float seed) {
  __shared__ float4 tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = make_float4(seed + i, 0.5f * i, 0.1f * i, 0.f);
  __syncthreads();
  const unsigned long long x0 = pk(seed + threadIdx.x * 1e-3f, seed + 0.5f), x1 = pk(seed * 2, seed * 3 + threadIdx.x);
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    float v[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = tab[(i * 8 + j) & 1023];
      unsigned long long a = fma2(x0, pk(w.x, w.x), pk(w.z, w.z));
      a = fma2(x1, pk(w.y, w.y), a);
      float a0, a1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
      v[0][j] = __cosf(a0);
      v[1][j] = __cosf(a1);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const unsigned p1a = cvt2(v[0][2 * c + 1], v[0][2 * c]), p1b = cvt2(v[1][2 * c + 1], v[1][2 * c]);
      const unsigned long long ra = fma2(pk(__uint_as_float(p1a << 16), __uint_as_float(p1b << 16)), pk(-1.f, -1.f),
                                         pk(v[0][2 * c], v[1][2 * c]));
      const unsigned long long rb = fma2(pk(__uint_as_float(p1a & 0xffff0000u), __uint_as_float(p1b & 0xffff0000u)),
                                         pk(-1.f, -1.f), pk(v[0][2 * c + 1], v[1][2 * c + 1]));
      float q0, q1, r0, r1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(ra));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(rb));
      acc ^= p1a ^ p1b ^ cvt2(r0, q0) ^ cvt2(r1, q1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

void run_mix(int threads) {
  float* out;
  cudaMalloc(&out, 148 * 2 * 1024 * sizeof(float));
  const int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kmix<<<148, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e0);
  kmix<<<148, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * threads * iters * 16;
  printf("%-28s %4d threads/SM: %.3f ms, %.2f cos/clk/SM at 1965 MHz\n", "producer mix (no sync)", threads, ms,
         ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

template <int MODE>
void run(const char* name, int threads) {
  float* out;
  cudaMalloc(&out, 148 * 2 * 1024 * sizeof(float));
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * threads * iters * 8;
  printf("%-28s %4d threads/SM: %.3f ms, %.2f ops/clk/SM at 1965 MHz\n", name, threads, ms,
         ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int th : {256, 512, 1024}) {
    run<0>("cos.approx (FMUL.RZ+MUFU)", th);
    run<1>("cos.approx + FFMA", th);
    run<2>("FMUL + MUFU.EX2", th);
    run<3>("MUFU.EX2 alone", th);
  }
  for (int th : {256, 512, 768, 1024}) run_mix(th);
  return 0;
}
