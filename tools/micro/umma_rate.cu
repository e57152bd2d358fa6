// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 in, fp32 acc, cta_group::1) for the operand layouts the
// BiDAF kernels use, no-swizzle core-matrix order.  One CTA, warp 0 issues REP MMAs back to back then commits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mmbidaf_b200/csrc tools/micro/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include "tc_common.cuh"
namespace mmb { void set_error(const char*, ...) {} }
using namespace mmb::tc;

struct Cfg { int m, n, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, b_mn; const char* name; };

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, int rep, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();
  if (warp_u == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
    const uint32_t a_lo = desc_lo(smem_u32(smem), c.a_lbo), b_lo = desc_lo(smem_u32(smem) + 100 * 1024, c.b_lbo);
    const uint32_t a_hi = (c.a_sbo >> 4) | (1u << 14), b_hi = (c.b_sbo >> 4) | (1u << 14);
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int r = 0; r < rep; ++r) {
        const int k = r & 3;
        umma_bf16_lh(tmem, a_lo + k * c.a_kstep / 16, a_hi, b_lo + k * c.b_kstep / 16, b_hi, idesc, r > 0, leader);
      }
      umma_commit(smem_u32(&bar), leader);
      const long long t1 = clock64();
      mbar_wait(smem_u32(&bar), pass);
      const long long t2 = clock64();
      if (leader && pass == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  const int G = 3328, G2 = 26 * 144;
  Cfg cfgs[] = {
      {128, 32, 128, G, 256, 128, G, 256, 0, "S-type   M128 N32  K-major x K-major (chunk stride 128)"},
      {128, 64, 128, G, 256, 128, G, 256, 0, "S-type   M128 N64  K-major x K-major"},
      {128, 128, 128, G, 256, 128, G, 256, 0, "S-type   M128 N128 K-major x K-major"},
      {128, 208, 128, G, 256, 128, G, 256, 0, "         M128 N208 K-major x K-major"},
      {64, 32, 128, G, 256, 128, G, 256, 0, "S-type   M64  N32  K-major x K-major"},
      {64, 64, 128, G, 256, 128, G, 256, 0, "S-type   M64  N64  K-major x K-major"},
      {128, 208, 2048, 128, 4096, G, 128, 2 * G, 1, "PV-type  M128 N208 P(K-major) x V(MN-major, chunk stride 128)"},
      {128, 208, 2048, 128, 4096, G2, 144, 2 * G2, 1, "PV-type  M128 N208 P(K-major) x V(MN-major, chunk stride 144)"},
      {128, 208, 2048, 128, 4096, 26 * 160, 160, 2 * 26 * 160, 1, "PV-type  M128 N208 P(K-major) x V(MN-major, chunk stride 160)"},
      {128, 208, 1024, 128, 2048, G, 128, 2 * G, 1, "PV-type  M128 N208 P(64-row compact) x V(MN-major, chunk stride 128)"},
      {128, 64, 2048, 128, 4096, G, 128, 2 * G, 1, "PV-type  M128 N64  P x V(MN-major, chunk stride 128)"},
      {128, 32, 144, G2, 288, 144, G2, 288, 0, "S-type   M128 N32  K-major x K-major (chunk stride 144)"},
  };
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rep = 64;
  for (const Cfg& c : cfgs) {
    rate_kernel<<<1, 128, 200 * 1024>>>(c, rep, out);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    printf("%-75s issue %6.1f  complete %6.1f cycles/MMA  (floor %d)\n", c.name, (double)h[0] / rep, (double)h[1] / rep, 128 * c.n / 256);
  }
  return 0;
}
