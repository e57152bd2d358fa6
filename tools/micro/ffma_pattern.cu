// Micro-benchmark: FFMA issue rate for the register pattern of the LSTM recurrence -- NW unique weight registers, 4 accumulator
// chains, the multiplicand re-used by 4 consecutive FFMAs -- against warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_pattern tools/micro/ffma_pattern.cu && /tmp/ffma_pattern
#include <cstdio>
#include <cuda_runtime.h>
template <int NW, int NACC>
__global__ void __launch_bounds__(256) k(float* out, const float* win, int iters, float seed) {
  float w[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = win[i * 32 + (threadIdx.x & 31)];
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = seed * i;
  float h[4] = {seed, seed * 2, seed * 3, seed * 4};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i % NACC] = fmaf(w[i], h[(i / NACC) & 3], acc[i % NACC]);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = h[i] * 0.999f;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NW, int NACC>
void run(float* out, float* win) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  for (int warps = 1; warps <= 8; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<NW, NACC><<<148, warps * 32>>>(out, win, iters, 1e-9f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per FFMA per scheduler (1.965 GHz): warps <= 4 -> one warp per scheduler
    const double per_sched = warps <= 4 ? 1.0 : warps / 4.0;
    printf("NW=%3d NACC=%d warps/CTA=%d: %.3f ms -> %.2f cycles per FFMA per scheduler\n", NW, NACC, warps, ms,
           ms * 1e-3 * 1.965e9 / ((double)iters * NW * per_sched));
  }
}
int main() {
  float *out, *win; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&win, 256 * 32 * 4); cudaMemset(win, 0, 256 * 32 * 4);
  run<8, 4>(out, win); run<8, 8>(out, win); run<64, 4>(out, win); run<208, 4>(out, win); run<208, 8>(out, win);
  return 0;
}
