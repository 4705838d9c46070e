// Which pipe do the BF16x3 producers' conversions share?  (profiling aid, round 2)
// The producer mix of the bench kernel (sp_tensor.cuh) tops out at ~12.6 cos/clk/SM with 1024 threads although
// neither the issue slots (6 of 8 per XU period) nor the XU pipe (16 MUFU/clk/SM) is full.  This probe times the
// same loop with different operand splits to see what F2FP (cvt.rn.bf16x2.f32) competes with.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mix_probe.bin scripts/mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

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
__device__ __forceinline__ unsigned prmt_hi(float lo, float hi) {  // {hi16(hi), hi16(lo)}: truncating bf16 pack
  unsigned r;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}

// MODE 0: the kernel's split (F2FP hi, IMAD.U32/LOP3 unpack, FFMA2 residual, F2FP lo)
// MODE 1: no split (projection + cos only)
// MODE 2: hi by truncation (PRMT pack, LOP3 mask), residual FFMA2, lo by F2FP
// MODE 3: hi and lo by truncation (PRMT, LOP3, FFMA2, PRMT): no F2FP at all
// MODE 4: MODE 0 without the LDS (table entries from registers)
// MODE 5: cos + one F2FP per pair only
// MODE 6: cos + one PRMT per pair only
template <int MODE>
__global__ void __launch_bounds__(1024) kmix(float* out, int iters, float seed) {
  __shared__ float4 tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = make_float4(seed + i, 0.5f * i, 0.1f * i, 0.f);
  __syncthreads();
  const unsigned long long x0 = pk(seed + threadIdx.x * 1e-3f, seed + 0.5f), x1 = pk(seed * 2, seed * 3 + threadIdx.x);
  unsigned acc = 0;
  float4 wreg = tab[threadIdx.x & 1023];
  for (int i = 0; i < iters; ++i) {
    float v[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 w;
      if (MODE == 4) { w = wreg; w.x += (float)j; }
      else w = tab[(i * 8 + j) & 1023];
      unsigned long long a = fma2(x0, pk(w.x, w.x), pk(w.z, w.z));
      a = fma2(x1, pk(w.y, w.y), a);
      float a0, a1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
      v[0][j] = __cosf(a0);
      v[1][j] = __cosf(a1);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (MODE == 0 || MODE == 4) {
        const unsigned p1a = cvt2(v[0][2 * c + 1], v[0][2 * c]), p1b = cvt2(v[1][2 * c + 1], v[1][2 * c]);
        const unsigned long long ra = fma2(pk(__uint_as_float(p1a << 16), __uint_as_float(p1b << 16)), pk(-1.f, -1.f),
                                           pk(v[0][2 * c], v[1][2 * c]));
        const unsigned long long rb = fma2(pk(__uint_as_float(p1a & 0xffff0000u), __uint_as_float(p1b & 0xffff0000u)),
                                           pk(-1.f, -1.f), pk(v[0][2 * c + 1], v[1][2 * c + 1]));
        float q0, q1, r0, r1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(ra));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(rb));
        acc ^= p1a ^ p1b ^ cvt2(r0, q0) ^ cvt2(r1, q1);
      } else if (MODE == 1) {
        acc ^= __float_as_uint(v[0][2 * c]) ^ __float_as_uint(v[0][2 * c + 1]) ^ __float_as_uint(v[1][2 * c]) ^
               __float_as_uint(v[1][2 * c + 1]);
      } else if (MODE == 2 || MODE == 3) {
        const unsigned p1a = prmt_hi(v[0][2 * c], v[0][2 * c + 1]), p1b = prmt_hi(v[1][2 * c], v[1][2 * c + 1]);
        const float h00 = __uint_as_float(__float_as_uint(v[0][2 * c]) & 0xffff0000u);
        const float h10 = __uint_as_float(__float_as_uint(v[1][2 * c]) & 0xffff0000u);
        const float h01 = __uint_as_float(__float_as_uint(v[0][2 * c + 1]) & 0xffff0000u);
        const float h11 = __uint_as_float(__float_as_uint(v[1][2 * c + 1]) & 0xffff0000u);
        const unsigned long long ra = fma2(pk(h00, h10), pk(-1.f, -1.f), pk(v[0][2 * c], v[1][2 * c]));
        const unsigned long long rb = fma2(pk(h01, h11), pk(-1.f, -1.f), pk(v[0][2 * c + 1], v[1][2 * c + 1]));
        float q0, q1, r0, r1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(ra));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(rb));
        if (MODE == 2) acc ^= p1a ^ p1b ^ cvt2(r0, q0) ^ cvt2(r1, q1);
        else acc ^= p1a ^ p1b ^ prmt_hi(q0, r0) ^ prmt_hi(q1, r1);
      } else if (MODE == 5) {
        acc ^= cvt2(v[0][2 * c + 1], v[0][2 * c]) ^ cvt2(v[1][2 * c + 1], v[1][2 * c]);
      } else if (MODE == 6) {
        acc ^= prmt_hi(v[0][2 * c], v[0][2 * c + 1]) ^ prmt_hi(v[1][2 * c], v[1][2 * c + 1]);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

// MUFU.COS with k extra F2FP (or PRMT) per cos, nothing else: do they share a pipe?
template <int NCVT, bool USE_PRMT>
__global__ void __launch_bounds__(1024) kcvt(float* out, int iters, float seed) {
  float a[8];
  unsigned acc = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x * 1e-3f + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = __cosf(a[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < NCVT; ++c) {
        const float o = a[(j + 1 + c) & 7];
        acc ^= USE_PRMT ? prmt_hi(a[j], o) : cvt2(a[j], o);
      }
  }
  float s = __uint_as_float(acc);
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void time_kernel(const char* name, K kern, int threads, int iters, double ops_per_thread_iter) {
  float* out;
  cudaMalloc(&out, 148 * 2 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<148, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e0);
  kern<<<148, threads>>>(out, iters, 0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * threads * iters * ops_per_thread_iter;
  printf("%-44s %4d thr/SM: %7.3f ms  %6.2f cos/clk/SM @1965MHz\n", name, threads, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int th : {512, 768, 1024}) {
    time_kernel("mix 0: kernel split (F2FP,F2FP)", kmix<0>, th, 2048, 16);
    time_kernel("mix 1: no split", kmix<1>, th, 2048, 16);
    time_kernel("mix 2: hi PRMT trunc, lo F2FP", kmix<2>, th, 2048, 16);
    time_kernel("mix 3: hi PRMT, lo PRMT (no F2FP)", kmix<3>, th, 2048, 16);
    time_kernel("mix 4: kernel split, no LDS", kmix<4>, th, 2048, 16);
    time_kernel("mix 5: cos + 0.5 F2FP", kmix<5>, th, 2048, 16);
    time_kernel("mix 6: cos + 0.5 PRMT", kmix<6>, th, 2048, 16);
  }
  for (int th : {512, 1024}) {
    time_kernel("cos + 1 F2FP", kcvt<1, false>, th, 4096, 8);
    time_kernel("cos + 2 F2FP", kcvt<2, false>, th, 4096, 8);
    time_kernel("cos + 4 F2FP", kcvt<4, false>, th, 4096, 8);
    time_kernel("cos + 1 PRMT", kcvt<1, true>, th, 4096, 8);
    time_kernel("cos + 2 PRMT", kcvt<2, true>, th, 4096, 8);
    time_kernel("cos + 4 PRMT", kcvt<4, true>, th, 4096, 8);
  }
  return 0;
}
