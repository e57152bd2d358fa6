// Micro-benchmark: stage the A operand of tcgen05.mma into TENSOR MEMORY with tcgen05.cp (shared memory -> TMEM, issued by the
// MMA thread, ordered with the MMAs in the tensor pipe) instead of ld.shared + tcgen05.st by all threads.
//   1. layout check: A tile (128 x 208 bf16, core-matrix order of tc_common.cuh) copied with 13 x tcgen05.cp.128x256b
//      (one per K step: 128 rows x 32 bytes -> 8 TMEM columns), described by the SAME no-swizzle K-major descriptor the MMA
//      would take for that K step (LBO 128, SBO GROUP_BYTES); product checked against the host.
//   2. cycles for the 13 copies + commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -I mmbidaf_b200/csrc tools/micro/tmem_cp.cu -o tools/micro/tmem_cp
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_common.cuh"
namespace mmb { void set_error(const char*, ...) {} }
using namespace mmb::tc;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t d_tmem, uint32_t lo, uint32_t hi, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .b64 ds;\n\t"
      "setp.ne.b32 q, %3, 0;\n\t"
      "mov.b64 ds, {%1, %2};\n\t"
      "@q tcgen05.cp.cta_group::1.128x256b [%0], ds;\n\t"
      "}" ::"r"(d_tmem), "r"(lo), "r"(hi), "r"(leader)
      : "memory");
}

constexpr int M = 128, K = DPAD, KSTEPS = K / 16;
constexpr int COL_D = 0, COL_A = 256;

__global__ void __launch_bounds__(128, 1) tmem_cp_kernel(const __nv_bfloat16* a_pack, const __nv_bfloat16* b_pack, float* d, int n,
                                                         long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  unsigned char* As = smem;
  unsigned char* Bs = smem + 64 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M / 8 * GROUP_BYTES / 16; i += 128) reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(a_pack)[i];
  for (int i = tid; i < n / 8 * GROUP_BYTES / 16; i += 128) reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(b_pack)[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();
  const uint32_t idesc = idesc_bf16(n, 0, 128);
  const uint32_t b_lo = desc_lo(smem_u32(Bs), 128), b_hi = desc_hi(GROUP_BYTES);
  const uint32_t a_lo = desc_lo(smem_u32(As), 128), a_hi = desc_hi(GROUP_BYTES);
  if (warp_u == 0) {
    uint32_t phase = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) tmem_cp_128x256b(tmem + COL_A + k * 8, a_lo + k * 16, a_hi, leader);
      umma_commit(smem_u32(&bar), leader);
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (leader) cycles[pass] = t1 - t0;
    }
    // the copies and the MMAs are ordered in the tensor pipe: no wait between them is needed; do it once more to use that
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) tmem_cp_128x256b(tmem + COL_A + k * 8, a_lo + k * 16, a_hi, leader);
    for (int k = 0; k < KSTEPS; ++k) umma_bf16_ts(tmem + COL_D, tmem + COL_A + k * 8, b_lo + k * 16, b_hi, idesc, k > 0, leader);
    umma_commit(smem_u32(&bar), leader);
    mbar_wait(smem_u32(&bar), phase);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int q = 0; q < n / 16; ++q) {
    float o[16];
    tmem_ld16(lane_base + COL_D + q * 16, o);
    for (int i = 0; i < 16; ++i) d[(size_t)tid * n + q * 16 + i] = o[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void pack(const std::vector<float>& src, int rows, std::vector<__nv_bfloat16>& dst) {
  dst.assign((size_t)rows / 8 * CHUNKS * 64, __float2bfloat16(0.f));
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k)
      dst[((size_t)(r / 8) * CHUNKS + k / 8) * 64 + (r % 8) * 8 + k % 8] = __float2bfloat16(src[(size_t)r * K + k]);
}

int main() {
  __nv_bfloat16 *da, *db;
  float* dd;
  long long* dc;
  cudaMalloc(&da, M / 8 * GROUP_BYTES);
  cudaMalloc(&db, 208 / 8 * GROUP_BYTES);
  cudaMalloc(&dd, M * 208 * 4);
  cudaMalloc(&dc, 16);
  cudaFuncSetAttribute(tmem_cp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int n : {32, 64}) {
    std::vector<float> a((size_t)M * K), b((size_t)n * K);
    srand(11 + n);
    for (auto& v : a) v = (float)(rand() % 17 - 8) / 8.f;
    for (auto& v : b) v = (float)(rand() % 13 - 6) / 4.f;
    std::vector<__nv_bfloat16> ap, bp;
    pack(a, M, ap);
    pack(b, n, bp);
    cudaMemcpy(da, ap.data(), ap.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice);
    tmem_cp_kernel<<<1, 128, 160 * 1024>>>(da, db, dd, n, dc);
    std::vector<float> d((size_t)M * n);
    long long cyc[2];
    cudaError_t e = cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)a[(size_t)i * K + k] * b[(size_t)j * K + k];
        err = fmax(err, fabs(s - d[(size_t)i * n + j]));
      }
    printf("tcgen05.cp 128x256b x13 -> A in TMEM, N%-3d: layout check max |err| %.3g (%s); 13 copies + commit: %lld cycles (second pass %lld)\n",
           n, err, err < 1e-3 ? "ok" : "LAYOUT WRONG", cyc[0], cyc[1]);
  }
  return 0;
}
