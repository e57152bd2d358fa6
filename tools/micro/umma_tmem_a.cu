// Micro-benchmark for the round-2 plan (profiles/r01_bidaf_tc_ncu.md): tcgen05.mma with the A operand in TENSOR MEMORY.
// The forward tile loop is bound by the tensor pipe re-reading the 4 KB A operand (the X tile, constant over the whole loop)
// from shared memory for every 128 x N x 16 instruction: 40 cycles at N = 32 against a floor of 16.  With A in TMEM
// (`tcgen05.mma ... [d], [a], b_desc, idesc, p`) only B comes from shared memory.  This program
//   1. checks the A-in-TMEM layout assumed below against a host product (row i of A = TMEM lane i, 32-bit column c of the
//      operand = K elements 2c, 2c+1 as a bf16 pair), for M = 128, K = 208, N = 32 .. 208, B K-major in core-matrix order;
//   2. prints cycles per MMA for A from TMEM next to A from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -I mmbidaf_b200/csrc tools/micro/umma_tmem_a.cu -o tools/micro/umma_tmem_a
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc_common.cuh"
namespace mmb { void set_error(const char*, ...) {} }
using namespace mmb::tc;

// D[tmem] (+)= A[tmem] * B[smem]
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

constexpr int M = 128, K = DPAD, KSTEPS = K / 16;
constexpr int COL_D = 0, COL_A = 256;            // accumulator (<= 208 columns), A operand (K / 2 = 104 columns)

// a_pack / b_pack: core-matrix order [row/8][chunk][row%8][8 bf16] (tc_common.cuh); d: (M, n) fp32 row-major
// m = 128: row i of A and D on lane i.  m = 64: hypothesis under test -- A uses the lanes the accumulator uses
// (tools/micro/umma_m64_layout.cu: row i on lane (i / 16) * 32 + i % 16, the upper half of each 32-lane quarter idle).
__global__ void __launch_bounds__(128, 1) tmem_a_kernel(const __nv_bfloat16* a_pack, const __nv_bfloat16* b_pack, float* d, int n,
                                                        int rep, long long* cycles, int m) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  unsigned char* As = smem;                               // 16 groups x 3328 B
  unsigned char* Bs = smem + 64 * 1024;                   // n / 8 groups
  const int tid = threadIdx.x, warp = tid >> 5;
  const int arow = m == 128 ? tid : ((tid & 31) < 16 ? (tid >> 5) * 16 + (tid & 31) : -1);      // the A / D row this lane holds
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
  // stage A into TMEM: thread = row, 32-bit column c of the operand <- bf16 pair (K = 2c, 2c + 1)
  {
    const int sr = arow < 0 ? 0 : arow;
    const unsigned char* arow_p = As + (sr >> 3) * GROUP_BYTES + (sr & 7) * 16;     // chunk ch of this row at + ch * 128
    for (int q = 0; q < K / 32; ++q) {                                              // 16 columns = 32 K elements = 4 chunks
      float v[16];
      for (int ch = 0; ch < 4; ++ch) {
        const uint4 u = *reinterpret_cast<const uint4*>(arow_p + (q * 4 + ch) * 128);
        v[ch * 4 + 0] = __uint_as_float(u.x); v[ch * 4 + 1] = __uint_as_float(u.y);
        v[ch * 4 + 2] = __uint_as_float(u.z); v[ch * 4 + 3] = __uint_as_float(u.w);
      }
      tmem_st16(lane_base + COL_A + q * 16, v);
    }
    {                                                                               // K = 208: the last 16 elements (8 columns)
      float v[16] = {};
      for (int ch = 0; ch < 2; ++ch) {
        const uint4 u = *reinterpret_cast<const uint4*>(arow_p + ((K / 32) * 4 + ch) * 128);
        v[ch * 4 + 0] = __uint_as_float(u.x); v[ch * 4 + 1] = __uint_as_float(u.y);
        v[ch * 4 + 2] = __uint_as_float(u.z); v[ch * 4 + 3] = __uint_as_float(u.w);
      }
      tmem_st16(lane_base + COL_A + (K / 32) * 16, v);                              // columns 96..111 (104..111 unused zeros)
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();
  const uint32_t idesc = idesc_bf16(n, 0, m);
  const uint32_t b_lo = desc_lo(smem_u32(Bs), 128), b_hi = desc_hi(GROUP_BYTES);
  const uint32_t a_lo = desc_lo(smem_u32(As), 128), a_hi = desc_hi(GROUP_BYTES);
  if (warp_u == 0) {
    // (1) one product for the layout check
    for (int k = 0; k < KSTEPS; ++k) umma_bf16_ts(tmem + COL_D, tmem + COL_A + k * 8, b_lo + k * 16, b_hi, idesc, k > 0, leader);
    umma_commit(smem_u32(&bar), leader);
    mbar_wait(smem_u32(&bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int q = 0; q < n / 16; ++q) {
    float o[16];
    tmem_ld16(lane_base + COL_D + q * 16, o);
    if (arow >= 0)
      for (int i = 0; i < 16; ++i) d[(size_t)arow * n + q * 16 + i] = o[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp_u == 0) {
    // (2) rates: A from TMEM, then A from shared memory, `rep` MMAs back to back each (two passes, the second is timed)
    uint32_t phase = 1;
    for (int mode = 0; mode < 2; ++mode)
      for (int pass = 0; pass < 2; ++pass) {
        const long long t0 = clock64();
        if (mode == 0) {                                   // (k = r & 7: no integer division in the issue loop)
#pragma unroll 8
          for (int r = 0; r < rep; ++r) umma_bf16_ts(tmem + COL_D, tmem + COL_A + (r & 7) * 8, b_lo + (r & 7) * 16, b_hi, idesc, 1, leader);
        } else {
#pragma unroll 8
          for (int r = 0; r < rep; ++r) umma_bf16_lh(tmem + COL_D, a_lo + (r & 7) * 16, a_hi, b_lo + (r & 7) * 16, b_hi, idesc, 1, leader);
        }
        umma_commit(smem_u32(&bar), leader);
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
        const long long t1 = clock64();
        if (leader && pass == 1) cycles[mode] = t1 - t0;
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void pack(const std::vector<float>& src, int rows, std::vector<__nv_bfloat16>& dst) {   // (rows, K) -> core-matrix order
  dst.assign((size_t)rows / 8 * CHUNKS * 64, __float2bfloat16(0.f));
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k)
      dst[((size_t)(r / 8) * CHUNKS + k / 8) * 64 + (r % 8) * 8 + k % 8] = __float2bfloat16(src[(size_t)r * K + k]);
}

int main() {
  const int rep = 104;
  __nv_bfloat16 *da, *db;
  float* dd;
  long long* dc;
  cudaMalloc(&da, M / 8 * GROUP_BYTES);
  cudaMalloc(&db, 208 / 8 * GROUP_BYTES);
  cudaMalloc(&dd, M * 208 * 4);
  cudaMalloc(&dc, 16);
  cudaFuncSetAttribute(tmem_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int m : {128, 64})
  for (int n : {32, 64, 128, 208}) {
    std::vector<float> a((size_t)M * K), b((size_t)n * K);
    srand(7 + n);
    for (auto& v : a) v = (float)(rand() % 17 - 8) / 8.f;          // exact in bf16: the check is exact
    for (auto& v : b) v = (float)(rand() % 13 - 6) / 4.f;
    std::vector<__nv_bfloat16> ap, bp;
    pack(a, M, ap);
    pack(b, n, bp);
    cudaMemcpy(da, ap.data(), ap.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice);
    tmem_a_kernel<<<1, 128, 160 * 1024>>>(da, db, dd, n, rep, dc, m);
    std::vector<float> d((size_t)M * n);
    long long cyc[2];
    cudaError_t e = cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)a[(size_t)i * K + k] * b[(size_t)j * K + k];
        err = fmax(err, fabs(s - d[(size_t)i * n + j]));
      }
    printf("M%-3d N%-3d K-major B: layout check max |err| %.3g (%s); A from TMEM %.1f cycles/MMA, A from smem %.1f (floor %d)\n", m, n, err,
           err < 1e-3 ? "ok" : "LAYOUT WRONG", (double)cyc[0] / rep, (double)cyc[1] / rep, 128 * n / 256);
  }
  return 0;
}
