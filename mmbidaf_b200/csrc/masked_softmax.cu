// masked_softmax over the last axis, forward and backward (layers/attention.py:78-98).
//   y = softmax(mask ? x : -1e30)       (or log_softmax)
// One CTA per row.  A fully masked row is uniform (all logits equal -1e30), exactly like the reference.
#include "common.cuh"

namespace mmb {
namespace {

constexpr int MS_THREADS = 128;

__device__ __forceinline__ float block_all_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = lane < MS_THREADS / 32 ? red[lane] : (is_max ? -INFINITY : 0.f);
  return is_max ? warp_max(r) : warp_sum(r);
}

__global__ void __launch_bounds__(MS_THREADS) masked_softmax_fwd_kernel(const float* __restrict__ x,
                                                                        const uint8_t* __restrict__ mask,
                                                                        float* __restrict__ y, int n, int log_mode) {
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const float* xr = x + row * n;
  const uint8_t* mr = mask + row * n;
  float* yr = y + row * n;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += MS_THREADS) mx = fmaxf(mx, mr[i] ? xr[i] : kNegFill);
  mx = block_all_reduce(mx, red, true);
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += MS_THREADS) sum += expf((mr[i] ? xr[i] : kNegFill) - mx);
  sum = block_all_reduce(sum, red, false);
  const float inv = 1.f / sum, lsum = logf(sum);
  for (int i = threadIdx.x; i < n; i += MS_THREADS) {
    const float z = (mr[i] ? xr[i] : kNegFill) - mx;
    yr[i] = log_mode ? z - lsum : expf(z) * inv;
  }
}

__global__ void __launch_bounds__(MS_THREADS) masked_softmax_bwd_kernel(const float* __restrict__ y,
                                                                        const float* __restrict__ dy,
                                                                        const uint8_t* __restrict__ mask,
                                                                        float* __restrict__ dx, int n, int log_mode) {
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const float* yr = y + row * n;
  const float* gr = dy + row * n;
  const uint8_t* mr = mask + row * n;
  float* dr = dx + row * n;
  float dot = 0.f;
  for (int i = threadIdx.x; i < n; i += MS_THREADS) dot += log_mode ? gr[i] : gr[i] * yr[i];
  dot = block_all_reduce(dot, red, false);
  for (int i = threadIdx.x; i < n; i += MS_THREADS) {
    const float g = log_mode ? gr[i] - expf(yr[i]) * dot : yr[i] * (gr[i] - dot);
    dr[i] = mr[i] ? g : 0.f;            // d(mask*x)/dx = mask
  }
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_masked_softmax_fwd(const float* x, const uint8_t* mask, float* y, long long rows, int n, int log_mode,
                                      mmb_stream_t stream) {
  MMB_REQUIRE(x && mask && y, MMB_ERR_INVALID, "mmb_masked_softmax_fwd: null pointer");
  MMB_REQUIRE(rows > 0 && n > 0 && rows < (1ll << 31), MMB_ERR_INVALID, "mmb_masked_softmax_fwd: rows=%lld n=%d", rows, n);
  mmb::masked_softmax_fwd_kernel<<<(unsigned)rows, mmb::MS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, mask, y, n,
                                                                                                          log_mode);
  return mmb::check_launch("masked_softmax_fwd_kernel");
}

extern "C" int mmb_masked_softmax_bwd(const float* y, const float* dy, const uint8_t* mask, float* dx, long long rows,
                                      int n, int log_mode, mmb_stream_t stream) {
  MMB_REQUIRE(y && dy && mask && dx, MMB_ERR_INVALID, "mmb_masked_softmax_bwd: null pointer");
  MMB_REQUIRE(rows > 0 && n > 0 && rows < (1ll << 31), MMB_ERR_INVALID, "mmb_masked_softmax_bwd: rows=%lld n=%d", rows, n);
  mmb::masked_softmax_bwd_kernel<<<(unsigned)rows, mmb::MS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(y, dy, mask, dx,
                                                                                                          n, log_mode);
  return mmb::check_launch("masked_softmax_bwd_kernel");
}
