// Probe: where do the rows of a tcgen05.mma cta_group::1 M=64 accumulator land in TMEM?
// A (64 x 16, K-major) has A[i][0] = i + 1, B (8 x 16) has B[n][0] = n + 1 -> D[i][n] = (i + 1)(n + 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mmbidaf_b200/csrc tools/micro/umma_m64_layout.cu -o tools/micro/umma_m64_layout
#include <cstdio>
#include "tc_common.cuh"
namespace mmb { void set_error(const char*, ...) {} }
using namespace mmb::tc;

__global__ void __launch_bounds__(128, 1) probe(float* out, int m) {
  __shared__ __align__(128) __nv_bfloat16 A[128 * 16];     // core-matrix order: [row/8][chunk 0..1][row%8][8]
  __shared__ __align__(128) __nv_bfloat16 B[16 * 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 16; i += 128) A[i] = __float2bfloat16(0.f);
  for (int i = tid; i < 16 * 16; i += 128) B[i] = __float2bfloat16(0.f);
  __syncthreads();
  if (tid < 128) A[(tid / 8) * 128 + (tid % 8) * 8] = __float2bfloat16((float)(tid + 1));   // chunk 0, element 0 of row tid
  if (tid < 16) B[(tid / 8) * 128 + (tid % 8) * 8] = __float2bfloat16((float)(tid + 1));
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 32);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  // clear the accumulator columns first (all 128 lanes)
  float z[16];
  for (int i = 0; i < 16; ++i) z[i] = -1.f;
  tmem_st16(tmem + ((uint32_t)(warp * 32) << 16), z);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t leader = elect_one();
  if (uniform_warp_idx() == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    umma_bf16_lh(tmem, desc_lo(smem_u32(A), 128), desc_hi(256), desc_lo(smem_u32(B), 128), desc_hi(256), idesc, 0, leader);
    umma_commit(smem_u32(&bar), leader);
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  float v[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = v[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 16 * 4);
  for (int m : {128, 64}) {
    probe<<<1, 128>>>(out, m);
    float h[128 * 16];
    cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("M=%d: %s\n", m, cudaGetErrorString(e)); return 1; }
    printf("M=%d: lane -> (column 0 value = row + 1, column 1 value / column 0)\n", m);
    for (int l = 0; l < 128; ++l) printf("%s%3d:%4.0f/%3.1f", l % 8 == 0 ? "\n  " : "  ", l, h[l * 16], h[l * 16] != 0 ? h[l * 16 + 1] / h[l * 16] : 0.f);
    printf("\n");
  }
  return 0;
}
