// Issue cost of the producers' instruction kinds on sm_100 (profiling aid, round 2).
// One CTA of 512 threads per SM (4 warps per scheduler, like the bench kernel's producers); every thread runs
// straight-line code over 8-16 independent register chains.  Reports warp-instructions per clock per scheduler
// (1.0 = one issue slot per instruction) for each kind alone and for the pairs that matter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/issue_probe.bin scripts/issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

enum { K_FFMA, K_FFMA2, K_FMULRZ, K_MUFU, K_F2FP, K_FHFMA, K_LOP3, K_IMAD, K_PRMT, K_FFMA2_F2FP, K_FFMA2_MUFU, K_F2FP_MUFU,
       K_FHFMA_F2FP, K_FFMA_F2FP, K_FFMA2_FFMA, K_MIX, K_N };
const char* names[K_N] = {"FFMA", "FFMA2", "FMUL.RZ", "MUFU.COS (raw: ex2)", "F2FP.BF16 pack", "FHFMA.BF16", "LOP3", "IMAD.U32", "PRMT",
                          "FFMA2 + F2FP", "FFMA2 + MUFU", "F2FP + MUFU", "FHFMA + F2FP", "FFMA + F2FP", "FFMA2 + FFMA",
                          "mix: 2 FFMA2, 2 FMUL.RZ, 2 MUFU, 2 F2FP, 2 FHFMA"};
const int per_iter[K_N] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 16, 16, 16, 16, 16, 16, 10 * 4};

template <int KIND>
__global__ void __launch_bounds__(512) kprobe(float* out, int iters, float seed) {
  float f[8], g[8];
  uint64_t d[8];
  uint32_t u[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f[j] = seed + threadIdx.x * 1e-3f + j; g[j] = seed * j; d[j] = pk(f[j], g[j]); u[j] = threadIdx.x * 977u + j; }
  const float c1 = seed * 1.0001f, c2 = seed * 0.5f;
  const uint64_t dc = pk(c1, c2);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (KIND == K_FFMA) {
#define X(j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(c1), "f"(c2));
        REP8(X)
#undef X
      } else if (KIND == K_FFMA2) {
#define X(j) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[j]) : "l"(dc));
        REP8(X)
#undef X
      } else if (KIND == K_FMULRZ) {
#define X(j) asm volatile("mul.rz.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(c1));
        REP8(X)
#undef X
      } else if (KIND == K_MUFU) {
#define X(j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
        REP8(X)
#undef X
      } else if (KIND == K_F2FP) {
#define X(j) asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.bf16x2.f32 %0, t, %1;}" : "+r"(u[j]) : "f"(g[j]));
        REP8(X)
#undef X
      } else if (KIND == K_FHFMA) {
#define X(j) asm volatile("{.reg .b16 l, h, m; mov.b32 {l, h}, %1; mov.b16 m, 0xBF80; fma.rn.f32.bf16 %0, l, m, %0;}" : "+f"(f[j]) : "r"(u[j]));
        REP8(X)
#undef X
      } else if (KIND == K_LOP3) {
#define X(j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(u[(j + 1) & 7]), "r"(u[(j + 2) & 7]));
        REP8(X)
#undef X
      } else if (KIND == K_IMAD) {
#define X(j) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(u[(j + 1) & 7]), "r"(u[(j + 2) & 7]));
        REP8(X)
#undef X
      } else if (KIND == K_PRMT) {
#define X(j) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[j]) : "r"(u[(j + 1) & 7]));
        REP8(X)
#undef X
      } else if (KIND == K_FFMA2_F2FP) {
#define X(j) asm volatile("fma.rn.f32x2 %0, %0, %2, %2;\n\t{.reg .f32 t; mov.b32 t, %1; cvt.rn.bf16x2.f32 %1, t, %3;}" : "+l"(d[j]), "+r"(u[j]) : "l"(dc), "f"(g[j]));
        REP8(X)
#undef X
      } else if (KIND == K_FFMA2_MUFU) {
#define X(j) asm volatile("fma.rn.f32x2 %0, %0, %2, %2;\n\tex2.approx.ftz.f32 %1, %1;" : "+l"(d[j]), "+f"(f[j]) : "l"(dc));
        REP8(X)
#undef X
      } else if (KIND == K_F2FP_MUFU) {
#define X(j) asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.bf16x2.f32 %0, t, %2;}\n\tex2.approx.ftz.f32 %1, %1;" : "+r"(u[j]), "+f"(f[j]) : "f"(g[j]));
        REP8(X)
#undef X
      } else if (KIND == K_FHFMA_F2FP) {
#define X(j) asm volatile("{.reg .b16 l, h, m; mov.b32 {l, h}, %2; mov.b16 m, 0xBF80; fma.rn.f32.bf16 %0, l, m, %0;}\n\t{.reg .f32 t; mov.b32 t, %1; cvt.rn.bf16x2.f32 %1, t, %3;}" : "+f"(f[j]), "+r"(u[j]) : "r"(u[(j + 1) & 7]), "f"(g[j]));
        REP8(X)
#undef X
      } else if (KIND == K_FFMA_F2FP) {
#define X(j) asm volatile("fma.rn.f32 %0, %0, %2, %3;\n\t{.reg .f32 t; mov.b32 t, %1; cvt.rn.bf16x2.f32 %1, t, %4;}" : "+f"(f[j]), "+r"(u[j]) : "f"(c1), "f"(c2), "f"(g[j]));
        REP8(X)
#undef X
      } else if (KIND == K_FFMA2_FFMA) {
#define X(j) asm volatile("fma.rn.f32x2 %0, %0, %2, %2;\n\tfma.rn.f32 %1, %1, %3, %4;" : "+l"(d[j]), "+f"(f[j]) : "l"(dc), "f"(c1), "f"(c2));
        REP8(X)
#undef X
      } else if (KIND == K_MIX) {
        // per 4 evaluations (2 features x 2 rows): 2 FFMA2, 4 FMUL.RZ + 4 MUFU -> here 2+2 (half), kept proportional:
        // 2 FFMA2, 2 FMUL.RZ, 2 MUFU, 2 F2FP, 2 FHFMA per unit, 4 units
#define X(j) asm volatile("fma.rn.f32x2 %0, %0, %3, %3;\n\tmul.rz.f32 %1, %1, %4;\n\tex2.approx.ftz.f32 %1, %1;\n\t" \
                          "cvt.rn.bf16x2.f32 %2, %1, %4;\n\t{.reg .b16 l, h, m; mov.b32 {l, h}, %2; mov.b16 m, 0xBF80; fma.rn.f32.bf16 %1, l, m, %1;}" \
                          : "+l"(d[j]), "+f"(f[j]), "+r"(u[j]) : "l"(dc), "f"(c1));
        REP8(X)
#undef X
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d[j]));
    s += f[j] + g[j] + lo + hi + __uint_as_float(u[j]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(float* out) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kprobe<KIND><<<148, 512>>>(out, iters, 0.3f);
  cudaEventRecord(e0);
  kprobe<KIND><<<148, 512>>>(out, iters, 0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double n = KIND == K_MIX ? 5.0 * 8 * 4 : per_iter[KIND] * 4.0;  // warp-instructions per warp per iteration
  const double warp_inst = 148.0 * 16 * iters * n;
  const double clk = ms * 1e-3 * 1.965e9;
  printf("%-52s %7.3f ms  %5.2f warp-inst/clk/scheduler (at 1965 MHz)\n", names[KIND], ms, warp_inst / clk / 148 / 4);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  run<K_FFMA>(out); run<K_FFMA2>(out); run<K_FMULRZ>(out); run<K_MUFU>(out); run<K_F2FP>(out); run<K_FHFMA>(out);
  run<K_LOP3>(out); run<K_IMAD>(out); run<K_PRMT>(out); run<K_FFMA2_F2FP>(out); run<K_FFMA2_MUFU>(out); run<K_F2FP_MUFU>(out);
  run<K_FHFMA_F2FP>(out); run<K_FFMA_F2FP>(out); run<K_FFMA2_FFMA>(out); run<K_MIX>(out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
