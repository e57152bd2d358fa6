// Column sums of a tall row-major matrix: out[c] = sum_r a[r][c], a (n, p) fp32 with n up to B*L = 32768 rows and p a
// few hundred columns.  Every bias gradient of the step is one (LSTM: d(pre-activations) (B*L, 8H) summed over B*L,
// the gradient of b_ih + b_hh, layers/encoding.py:76-81; highway: encoding.py:52-59; the decoder's hoisted projections,
// attention.py:152-157).  HBM-bound: read a once.  Two deterministic stages (no atomics): a (column block x row block)
// grid sums float4 columns over its row slice into partial (R, p), a second launch adds the R partials in a fixed order.
#include "common.cuh"

namespace mmb {
namespace {

constexpr int CS_TX = 32, CS_TY = 8;            // 32 float4 = 128 columns per block, 8 rows in flight per iteration

__global__ void __launch_bounds__(CS_TX * CS_TY) col_sum_partial_kernel(const float* __restrict__ a, float* __restrict__ partial,
                                                                        long long n, int p, int rows_per_block) {
  __shared__ float4 red[CS_TY][CS_TX];
  const int tx = threadIdx.x % CS_TX, ty = threadIdx.x / CS_TX;
  const int c4 = blockIdx.x * CS_TX + tx;       // float4 column
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < p) {
    const float4* src = reinterpret_cast<const float4*>(a) + c4;
    const long long stride4 = p >> 2;
    long long r = r0 + ty;
    for (; r + 3 * CS_TY < r1; r += 4 * CS_TY) {          // four independent loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (r + u * CS_TY) * stride4);
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
    }
    for (; r < r1; r += CS_TY) {
      const float4 v = __ldg(src + r * stride4);
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
    }
  }
  float4 s = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                         (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c4 * 4 < p) {
#pragma unroll
    for (int y = 1; y < CS_TY; ++y) { const float4 o = red[y][tx]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
    reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * p)[c4] = s;
  }
}

// Same row-block partial sums for any p / alignment (one float column per thread): the shapes the float4 kernel does not take
// (p % 4 != 0 only happens for toy hidden sizes; the result goes through the same fixed-order final sum).
__global__ void __launch_bounds__(CS_TX * CS_TY) col_sum_partial_scalar_kernel(const float* __restrict__ a, float* __restrict__ partial,
                                                                               long long n, int p, int rows_per_block) {
  __shared__ float red[CS_TY][CS_TX];
  const int tx = threadIdx.x % CS_TX, ty = threadIdx.x / CS_TX;
  const int c = blockIdx.x * CS_TX + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  float acc = 0.f;
  if (c < p)
    for (long long r = r0 + ty; r < r1; r += CS_TY) acc += __ldg(a + r * p + c);
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < p) {
#pragma unroll
    for (int y = 1; y < CS_TY; ++y) acc += red[y][tx];
    partial[(size_t)blockIdx.y * p + c] = acc;
  }
}

// out[c] = sum_r partial[r][c]: 32 columns x 8 row lanes per block (a column's R partials are summed by 8 threads with
// independent loads, then across the lanes in a fixed order) -- one thread per column walked its R rows serially in ~11 us.
__global__ void __launch_bounds__(256) col_sum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int p, int R) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f;
  if (c < p) {
    int r = ty;
    for (; r + 8 < R; r += 16) { s0 += partial[(size_t)r * p + c]; s1 += partial[(size_t)(r + 8) * p + c]; }
    if (r < R) s0 += partial[(size_t)r * p + c];
  }
  red[ty][tx] = s0 + s1;
  __syncthreads();
  if (ty == 0 && c < p) {
    float s = red[0][tx];
#pragma unroll
    for (int y = 1; y < 8; ++y) s += red[y][tx];
    out[c] = s;
  }
}

}  // namespace
}  // namespace mmb

// Row blocks (= rows of the `partial` workspace) for an (n, p) matrix: about four blocks per SM in total, >= 64 rows each.
extern "C" int mmb_col_sum_blocks(long long n, int p) {
  if (n <= 0 || p <= 0) return 0;
  const int col_blocks = (p + 4 * mmb::CS_TX - 1) / (4 * mmb::CS_TX);
  long long R = (4 * 148 + col_blocks - 1) / col_blocks;
  const long long max_r = (n + 63) / 64;
  if (R > max_r) R = max_r;
  return (int)(R < 1 ? 1 : R);
}

extern "C" int mmb_col_sum(const float* a, float* partial, float* out, long long n, int p, mmb_stream_t stream) {
  MMB_REQUIRE(a && partial && out && n > 0 && p > 0, MMB_ERR_INVALID, "mmb_col_sum: bad arguments");
  const bool vec = p % 4 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(partial) & 15) == 0;
  const int R = mmb_col_sum_blocks(n, p);
  const int rows_per_block = (int)((n + R - 1) / R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) {
    const int col_blocks = (p + 4 * mmb::CS_TX - 1) / (4 * mmb::CS_TX);
    mmb::col_sum_partial_kernel<<<dim3(col_blocks, R), mmb::CS_TX * mmb::CS_TY, 0, st>>>(a, partial, n, p, rows_per_block);
  } else {
    const int col_blocks = (p + mmb::CS_TX - 1) / mmb::CS_TX;
    mmb::col_sum_partial_scalar_kernel<<<dim3(col_blocks, R), mmb::CS_TX * mmb::CS_TY, 0, st>>>(a, partial, n, p, rows_per_block);
  }
  if (int rc = mmb::check_launch("col_sum_partial_kernel")) return rc;
  mmb::col_sum_final_kernel<<<(p + 31) / 32, 256, 0, st>>>(partial, out, p, R);
  return mmb::check_launch("col_sum_final_kernel");
}
