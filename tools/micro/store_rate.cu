// Micro-benchmark: per-SM store throughput for the two epilogue store patterns of the BiDAF kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/store_rate.cu -o tools/micro/store_rate
// One CTA of 256 threads per SM writes ROWS x 800-byte blocks of (ROWS, 3200-byte) rows.
//   pattern 0: a warp instruction writes 512 contiguous bytes of ONE row
//   pattern 1: a warp instruction writes 64 bytes of each of 8 rows (the current epilogue)
//   pattern 2: like 0 but through cp.async.bulk (TMA) from shared memory, one 800-byte run per row
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROW_FLOATS = 800, BLK_FLOATS = 200;

__global__ void __launch_bounds__(256) store_kernel(float* out, int rows_per_cta, int pattern, int nblocks_of_row, long long* cycles) {
  extern __shared__ __align__(128) float stg[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* base = out + (size_t)blockIdx.x * rows_per_cta * ROW_FLOATS;
  for (int i = tid; i < 64 * BLK_FLOATS; i += 256) stg[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  const float4 v = make_float4(tid, 1.f, 2.f, 3.f);
  if (pattern == 0) {
    for (int blk = 0; blk < nblocks_of_row; ++blk)
      for (int r = warp; r < rows_per_cta; r += 8)
        for (int c4 = lane; c4 < BLK_FLOATS / 4; c4 += 32)
          *reinterpret_cast<float4*>(base + (size_t)r * ROW_FLOATS + blk * BLK_FLOATS + c4 * 4) = v;
  } else if (pattern == 1) {
    for (int blk = 0; blk < nblocks_of_row; ++blk)
      for (int it = warp; it < (rows_per_cta / 8) * 13; it += 8) {
        const int g8 = it / 13, b16 = it - g8 * 13;
        const int r = g8 * 8 + (lane & 7), col = b16 * 16 + (lane >> 3) * 4;
        if (col < BLK_FLOATS) *reinterpret_cast<float4*>(base + (size_t)r * ROW_FLOATS + blk * BLK_FLOATS + col) = v;
      }
  } else {
    // TMA bulk stores: 64 rows at a time from the staging buffer (contents irrelevant)
    for (int blk = 0; blk < nblocks_of_row; ++blk)
      for (int r0 = 0; r0 < rows_per_cta; r0 += 64) {
        if (tid < 64) {
          const uint32_t src = (uint32_t)__cvta_generic_to_shared(stg + tid * BLK_FLOATS);
          float* dst = base + (size_t)(r0 + tid) * ROW_FLOATS + blk * BLK_FLOATS;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(BLK_FLOATS * 4) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
      }
    if (tid < 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const int nsm = 148, rows = 128, reps = 3;      // 128 rows x 3 blocks x 800 B = 307 KB per CTA, as C2Q
  float* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)nsm * 4 * rows * ROW_FLOATS * 4);
  cudaMalloc(&cyc, nsm * 4 * 8);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * BLK_FLOATS * 4);
  for (int ctas : {148, 74, 296})
    for (int pattern = 0; pattern < 3; ++pattern) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      store_kernel<<<ctas, 256, 64 * BLK_FLOATS * 4>>>(out, rows, pattern, reps, cyc);
      cudaEventRecord(e0);
      store_kernel<<<ctas, 256, 64 * BLK_FLOATS * 4>>>(out, rows, pattern, reps, cyc);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long h[296];
      cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < ctas; ++i) mean += h[i];
      mean /= ctas;
      const double bytes = (double)rows * reps * BLK_FLOATS * 4;
      printf("ctas %3d pattern %d: %.0f cycles per CTA -> %.1f B/clk/CTA; kernel %.1f us -> %.2f TB/s\n", ctas, pattern, mean,
             bytes / mean, ms * 1e3, bytes * ctas / (ms * 1e-3) / 1e12);
    }
  return 0;
}
