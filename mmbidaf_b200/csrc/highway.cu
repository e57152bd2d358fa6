// Highway layer point-wise parts (layers/encoding.py:52-59): y = g * relu(t) + (1 - g) * x with
// g = sigmoid(pre[:, :H]), t = pre[:, H:], where pre = x [W_gate; W_transform]^T + bias is ONE plain GEMM
// done by the caller (the reference issues two GEMMs and seven element-wise kernels per layer).
#include "common.cuh"

namespace mmb {
namespace {

__global__ void __launch_bounds__(256) highway_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ x,
                                                          float* __restrict__ y, long long n, int H) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n * H) return;
  const long long r = i / H;
  const int c = (int)(i - r * H);
  const float g = sigmoidf_acc(pre[r * 2 * H + c]);
  const float t = fmaxf(pre[r * 2 * H + H + c], 0.f);
  const float xv = x[i];
  y[i] = g * t + (1.f - g) * xv;                      // encoding.py:57
}

// d_pre (n, 2H) = [d gate pre-activation | d transform pre-activation];  dx_direct = dy * (1 - g)
__global__ void __launch_bounds__(256) highway_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ x,
                                                          const float* __restrict__ dy, float* __restrict__ d_pre,
                                                          float* __restrict__ dx_direct, long long n, int H) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n * H) return;
  const long long r = i / H;
  const int c = (int)(i - r * H);
  const float pt = pre[r * 2 * H + H + c];
  const float g = sigmoidf_acc(pre[r * 2 * H + c]);
  const float t = fmaxf(pt, 0.f);
  const float go = dy[i];
  d_pre[r * 2 * H + c] = go * (t - x[i]) * g * (1.f - g);
  d_pre[r * 2 * H + H + c] = pt > 0.f ? go * g : 0.f;
  dx_direct[i] = go * (1.f - g);
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_highway_fwd(const float* pre, const float* x, float* y, long long n, int H, mmb_stream_t stream) {
  MMB_REQUIRE(pre && x && y && n > 0 && H > 0, MMB_ERR_INVALID, "mmb_highway_fwd: bad arguments");
  const long long blocks = (n * H + 255) / 256;
  MMB_REQUIRE(blocks < (1ll << 31), MMB_ERR_UNSUPPORTED, "mmb_highway_fwd: tensor too large");
  mmb::highway_fwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(pre, x, y, n, H);
  return mmb::check_launch("highway_fwd_kernel");
}

extern "C" int mmb_highway_bwd(const float* pre, const float* x, const float* dy, float* d_pre, float* dx_direct,
                               long long n, int H, mmb_stream_t stream) {
  MMB_REQUIRE(pre && x && dy && d_pre && dx_direct && n > 0 && H > 0, MMB_ERR_INVALID, "mmb_highway_bwd: bad arguments");
  const long long blocks = (n * H + 255) / 256;
  MMB_REQUIRE(blocks < (1ll << 31), MMB_ERR_UNSUPPORTED, "mmb_highway_bwd: tensor too large");
  mmb::highway_bwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(pre, x, dy, d_pre, dx_direct, n, H);
  return mmb::check_launch("highway_bwd_kernel");
}
