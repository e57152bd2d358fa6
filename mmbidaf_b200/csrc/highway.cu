// Highway layer point-wise parts (layers/encoding.py:52-59): y = g * relu(t) + (1 - g) * x with
// g = sigmoid(pre[:, :H]), t = pre[:, H:], where pre = x [W_gate; W_transform]^T + bias is ONE plain GEMM
// done by the caller (the reference issues two GEMMs and seven element-wise kernels per layer).
#include "common.cuh"

namespace mmb {
namespace {

// One thread = VEC consecutive columns of one row (VEC = 4 when H % 4 == 0: 128-bit accesses), a grid-stride loop with 32-bit index
// arithmetic.  (Round 1: one element per thread with a 64-bit division: 32 us for the 52 MB of the audio branch's layer, 1.6 TB/s, on
// the critical path of the step twice forward and twice backward.)
template <int VEC>
__global__ void __launch_bounds__(256) highway_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ x,
                                                          float* __restrict__ y, const unsigned n, const unsigned H) {
  const unsigned hv = H / VEC, total = n * hv;
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const unsigned r = i / hv, c = (i - r * hv) * VEC;
    const float* pr = pre + (size_t)r * 2 * H + c;
    float g[VEC], t[VEC], xv[VEC], o[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(g) = __ldg(reinterpret_cast<const float4*>(pr));
      *reinterpret_cast<float4*>(t) = __ldg(reinterpret_cast<const float4*>(pr + H));
      *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * H + c));
    } else {
      g[0] = pr[0]; t[0] = pr[H]; xv[0] = x[(size_t)r * H + c];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float gg = sigmoidf_acc(g[e]);
      o[e] = gg * fmaxf(t[e], 0.f) + (1.f - gg) * xv[e];             // encoding.py:57
    }
    if (VEC == 4) *reinterpret_cast<float4*>(y + (size_t)r * H + c) = *reinterpret_cast<float4*>(o);
    else y[(size_t)r * H + c] = o[0];
  }
}

// d_pre (n, 2H) = [d gate pre-activation | d transform pre-activation];  dx_direct = dy * (1 - g)
template <int VEC>
__global__ void __launch_bounds__(256) highway_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ x,
                                                          const float* __restrict__ dy, float* __restrict__ d_pre,
                                                          float* __restrict__ dx_direct, const unsigned n, const unsigned H) {
  const unsigned hv = H / VEC, total = n * hv;
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const unsigned r = i / hv, c = (i - r * hv) * VEC;
    const float* pr = pre + (size_t)r * 2 * H + c;
    float* dp = d_pre + (size_t)r * 2 * H + c;
    float gp[VEC], pt[VEC], xv[VEC], go[VEC], dg[VEC], dt[VEC], dx[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(gp) = __ldg(reinterpret_cast<const float4*>(pr));
      *reinterpret_cast<float4*>(pt) = __ldg(reinterpret_cast<const float4*>(pr + H));
      *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * H + c));
      *reinterpret_cast<float4*>(go) = __ldg(reinterpret_cast<const float4*>(dy + (size_t)r * H + c));
    } else {
      gp[0] = pr[0]; pt[0] = pr[H]; xv[0] = x[(size_t)r * H + c]; go[0] = dy[(size_t)r * H + c];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float g = sigmoidf_acc(gp[e]);
      const float t = fmaxf(pt[e], 0.f);
      dg[e] = go[e] * (t - xv[e]) * g * (1.f - g);
      dt[e] = pt[e] > 0.f ? go[e] * g : 0.f;
      dx[e] = go[e] * (1.f - g);
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(dp) = *reinterpret_cast<float4*>(dg);
      *reinterpret_cast<float4*>(dp + H) = *reinterpret_cast<float4*>(dt);
      *reinterpret_cast<float4*>(dx_direct + (size_t)r * H + c) = *reinterpret_cast<float4*>(dx);
    } else {
      dp[0] = dg[0]; dp[H] = dt[0]; dx_direct[(size_t)r * H + c] = dx[0];
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace mmb

extern "C" int mmb_highway_fwd(const float* pre, const float* x, float* y, long long n, int H, mmb_stream_t stream) {
  MMB_REQUIRE(pre && x && y && n > 0 && H > 0, MMB_ERR_INVALID, "mmb_highway_fwd: bad arguments");
  MMB_REQUIRE(n * H < (1ll << 31), MMB_ERR_UNSUPPORTED, "mmb_highway_fwd: tensor too large");
  const bool v4 = H % 4 == 0 && mmb::aligned16(pre) && mmb::aligned16(x) && mmb::aligned16(y);
  const long long work = v4 ? n * H / 4 : n * H;
  const unsigned blocks = (unsigned)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (v4) mmb::highway_fwd_kernel<4><<<blocks, 256, 0, st>>>(pre, x, y, (unsigned)n, (unsigned)H);
  else mmb::highway_fwd_kernel<1><<<blocks, 256, 0, st>>>(pre, x, y, (unsigned)n, (unsigned)H);
  return mmb::check_launch("highway_fwd_kernel");
}

extern "C" int mmb_highway_bwd(const float* pre, const float* x, const float* dy, float* d_pre, float* dx_direct,
                               long long n, int H, mmb_stream_t stream) {
  MMB_REQUIRE(pre && x && dy && d_pre && dx_direct && n > 0 && H > 0, MMB_ERR_INVALID, "mmb_highway_bwd: bad arguments");
  MMB_REQUIRE(n * H < (1ll << 31), MMB_ERR_UNSUPPORTED, "mmb_highway_bwd: tensor too large");
  const bool v4 = H % 4 == 0 && mmb::aligned16(pre) && mmb::aligned16(x) && mmb::aligned16(dy) && mmb::aligned16(d_pre) &&
                  mmb::aligned16(dx_direct);
  const long long work = v4 ? n * H / 4 : n * H;
  const unsigned blocks = (unsigned)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (v4) mmb::highway_bwd_kernel<4><<<blocks, 256, 0, st>>>(pre, x, dy, d_pre, dx_direct, (unsigned)n, (unsigned)H);
  else mmb::highway_bwd_kernel<1><<<blocks, 256, 0, st>>>(pre, x, dy, d_pre, dx_direct, (unsigned)n, (unsigned)H);
  return mmb::check_launch("highway_bwd_kernel");
}
