// Micro-benchmark for a round-2 idea (profiles/r01_step_timeline.md): the recurrent product of the LSTM step on the tensor
// cores, with W_hh RESIDENT IN TENSOR MEMORY as the A operand.  Today (csrc/bilstm.cu) a step is ~1300 cycles: a 208-FMA
// register mat-vec per thread, shuffles, activations, one barrier; the 1024 + 2 x 409 serial steps (forward and backward) are
// ~47 % of the training step.  This program times the serial chain such a step would have, nothing else:
//     h (N sequences x K = 112, bf16, shared memory, K-major core-matrix order)
//     -> 4 gate tiles x 7 K-steps = 28 tcgen05.mma 128 x N x 16 with A = W_hh tile in TMEM (4 x 56 columns)
//     -> commit -> mbarrier wait -> 4 x tcgen05.ld (this thread's unit: i, f, g, o for N sequences)
//     -> cell update (real activations) -> new h as bf16 into shared memory -> fence.proxy.async -> __syncthreads
// It prints cycles per step for N = 16 (the minimum N for M = 128).  No claim of numerical parity: the point is the latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -I mmbidaf_b200/csrc tools/micro/lstm_umma_step.cu -o tools/micro/lstm_umma_step
#include <cstdio>
#include "tc_common.cuh"
namespace mmb { void set_error(const char*, ...) {} }
using namespace mmb;
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

constexpr int H = 100, KP = 112, KSTEPS = KP / 16, KCH = KP / 8, N = 16;       // K padded to 112: 14 chunks, 56 TMEM columns
constexpr int HGROUP = KCH * 128;                                              // 8 sequences x 112 bf16
constexpr int COL_D = 0, COL_W = 64;                                           // D: 4 tiles x N columns; W: 4 tiles x 56 columns

__global__ void __launch_bounds__(128, 1) lstm_step_kernel(int steps, long long* cycles, float* sink) {
  __shared__ __align__(128) unsigned char hs[2][N / 8 * HGROUP];               // double-buffered h (bf16)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  for (int i = tid; i < 2 * N / 8 * HGROUP / 4; i += 128) reinterpret_cast<uint32_t*>(hs)[i] = 0x3c003c00u;   // h = 0.0078 (bf16 pairs)
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // W_hh into TMEM: thread = gate row of each tile (unit tid, gate m), bf16 pairs along K; small values so that c, h stay bounded
  for (int m = 0; m < 4; ++m)
    for (int q = 0; q < 4; ++q) {                                              // 56 columns = 3 x 16 + 8: write 64, the tail is padding
      float v[16];
      for (int i = 0; i < 16; ++i) {
        const __nv_bfloat162 w = __floats2bfloat162_rn(0.01f * ((tid + m + i) % 7 - 3), 0.01f * ((tid + q + i) % 5 - 2));
        v[i] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&w));
      }
      tmem_st16(lane_base + COL_W + m * 56 + q * 16, v);                       // (q = 3 spills 8 columns into the next tile, rewritten there)
    }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp_u = uniform_warp_idx();
  const uint32_t leader = elect_one();
  const uint32_t idesc = idesc_bf16(N, 0);
  float c[N], hsum = 0.f;
  for (int s = 0; s < N; ++s) c[s] = 0.f;
  uint32_t phase = 0;
  long long t0 = 0;
  for (int step = 0; step < steps + 8; ++step) {
    if (step == 8) t0 = clock64();                                             // 8 warm-up steps
    const int cur = step & 1;
    if (warp_u == 0) {
      const uint32_t b_lo = desc_lo(smem_u32(hs[cur]), 128), b_hi = desc_hi(HGROUP);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k)
          umma_bf16_ts(tmem + COL_D + m * N, tmem + COL_W + m * 56 + k * 8, b_lo + k * 16, b_hi, idesc, k > 0, leader);
      umma_commit(smem_u32(&bar), leader);
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    tc_fence_after();
    float g[4][N];
#pragma unroll
    for (int m = 0; m < 4; ++m) tmem_ld16(lane_base + COL_D + m * N, g[m]);
    // cell update of unit `tid` for the N sequences (gate order i, f, g, o), new h as bf16 into the other buffer
    unsigned char* hn = hs[cur ^ 1];
#pragma unroll
    for (int s = 0; s < N; ++s) {
      const float ig = gate_act(g[0][s], 1.f), fg = gate_act(g[1][s], 1.f), gg = gate_act(g[2][s], 2.f), og = gate_act(g[3][s], 1.f);
      c[s] = fg * c[s] + ig * gg;
      const float h = og * tanh_fast(c[s]);
      hsum += h;
      if (tid < H)
        *reinterpret_cast<__nv_bfloat16*>(hn + (s >> 3) * HGROUP + (tid >> 3) * 128 + (s & 7) * 16 + (tid & 7) * 2) = __float2bfloat16_rn(h);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  sink[tid] = hsum;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dc;
  float* ds;
  cudaMalloc(&dc, 8);
  cudaMalloc(&ds, 128 * 4);
  const int steps = 1024;
  lstm_step_kernel<<<1, 128>>>(steps, dc, ds);
  long long cyc;
  cudaError_t e = cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  printf("LSTM step on tcgen05 (W_hh in TMEM, H = 100, %d sequences per CTA, 128 threads): %.0f cycles per step over %d steps "
         "(register-resident FFMA kernel today: ~1300 cycles per step, one sequence per CTA)\n", N, (double)cyc / steps, steps);
  return 0;
}
