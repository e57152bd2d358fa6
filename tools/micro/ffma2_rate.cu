// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_rate tools/micro/ffma2_rate.cu && /tmp/ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8], w[8];
  unsigned long long a2[8], w2[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = seed * i; w[i] = 1.0f + seed * (i + threadIdx.x);
    float2 t = make_float2(a[i], a[i] + 1), u = make_float2(w[i], w[i] * 0.5f);
    a2[i] = *reinterpret_cast<unsigned long long*>(&t); w2[i] = *reinterpret_cast<unsigned long long*>(&u);
  }
  float h = seed + threadIdx.x;
  float2 hh = make_float2(h, h * 0.25f);
  unsigned long long h2 = *reinterpret_cast<unsigned long long*>(&hh);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(w[i]), "f"(h));
        else a2[i] = ffma2(w2[i], h2, a2[i]);
      }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&a2[i]); s += a[i] + t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 1; warps <= 8; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1e-9f); else k<1><<<148, warps * 32>>>(out, iters, 1e-9f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = (double)iters * 128, fma_per_instr = mode ? 2 : 1;
      printf("%s warps/CTA=%d: %.3f ms, %.2f instr/ns/SM, %.1f TFMA/s chip (x2 = TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", warps, ms,
             instr * warps / (ms * 1e6), instr * warps * 32 * fma_per_instr * 148 / (ms * 1e-3) / 1e12);
    }
  return 0;
}
