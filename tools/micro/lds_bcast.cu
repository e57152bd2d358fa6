// Micro-benchmark: cost of broadcast LDS.128 when a warp reads 1, 2 (even / odd lanes) or 4 distinct addresses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/lds_bcast.cu -o tools/micro/lds_bcast
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k(int groups, int iters, float* out, long long* cyc) {
  __shared__ __align__(16) float buf[4 * 128];
  for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) buf[i] = i * 0.001f;
  __syncthreads();
  const int g = threadIdx.x % groups;                   // which of `groups` address streams this lane follows
  const float* p = buf + g * 128;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k4 = 0; k4 < 26; ++k4) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"((unsigned)__cvta_generic_to_shared(p + k4 * 4)));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int threads : {224, 256})
    for (int groups : {1, 2, 4}) {
      const int iters = 200;
      k<<<1, threads>>>(groups, iters, out, cyc);
      k<<<1, threads>>>(groups, iters, out, cyc);
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%d threads, %d address streams per warp: %.1f cycles per 26 LDS.128 per warp-set (%.2f cycles per LDS.128 instruction across the CTA)\n",
             threads, groups, (double)h / iters, (double)h / iters / 26);
    }
  return 0;
}
