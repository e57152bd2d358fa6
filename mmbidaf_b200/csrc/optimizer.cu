// The parameter update of the reference training step (train.py:154-155: clip_grad_norm_(max_grad_norm) then
// Adadelta(lr, rho 0.9, eps 1e-6, weight_decay), train.py:110) as ONE pass over the flat parameter / gradient / state buffers
// of trainer.py: the clip coefficient is formed from the gradient norm (a device scalar, so nothing syncs), the clipped
// gradient is written back (p.grad keeps torch's semantics) and the four Adadelta updates follow in registers.  As ten
// element-wise ATen passes this was 77 us of the 5.3 ms step; HBM-bound here: 4 reads + 4 writes of 12.8 MB.
#include "common.cuh"

namespace mmb {
namespace {

__global__ void __launch_bounds__(256) adadelta_clip_kernel(float4* __restrict__ param, float4* __restrict__ grad,
                                                            float4* __restrict__ square_avg, float4* __restrict__ acc_delta,
                                                            const float* __restrict__ grad_norm, float max_norm, float lr,
                                                            float rho, float eps, float wd, long long n4) {
  // torch.nn.utils.clip_grad_norm_: coef = min(max_norm / (norm + 1e-6), 1)
  const float coef = fminf(max_norm / (grad_norm[0] + 1e-6f), 1.0f);
  const float one_m_rho = 1.0f - rho;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 g4 = grad[i], p4 = param[i], s4 = square_avg[i], a4 = acc_delta[i];
    float* g = reinterpret_cast<float*>(&g4);
    float* p = reinterpret_cast<float*>(&p4);
    float* s = reinterpret_cast<float*>(&s4);
    float* a = reinterpret_cast<float*>(&a4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      g[e] *= coef;
      const float ge = wd != 0.f ? g[e] + wd * p[e] : g[e];
      s[e] = rho * s[e] + one_m_rho * ge * ge;                         // square_avg
      const float delta = sqrtf(a[e] + eps) / sqrtf(s[e] + eps) * ge;
      a[e] = rho * a[e] + one_m_rho * delta * delta;                   // acc_delta
      p[e] -= lr * delta;
    }
    grad[i] = g4;
    param[i] = p4;
    square_avg[i] = s4;
    acc_delta[i] = a4;
  }
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_adadelta_clip_step(float* param, float* grad, float* square_avg, float* acc_delta, const float* grad_norm,
                                      float max_norm, float lr, float rho, float eps, float weight_decay, long long n,
                                      mmb_stream_t stream) {
  MMB_REQUIRE(param && grad && square_avg && acc_delta && grad_norm && n > 0, MMB_ERR_INVALID, "mmb_adadelta_clip_step: bad arguments");
  const uintptr_t bits = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(square_avg) | reinterpret_cast<uintptr_t>(acc_delta);
  MMB_REQUIRE(n % 4 == 0 && (bits & 15) == 0, MMB_ERR_UNSUPPORTED,
              "mmb_adadelta_clip_step: n=%lld must be a multiple of 4 and the buffers 16-byte aligned", n);
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mmb::adadelta_clip_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(param), reinterpret_cast<float4*>(grad), reinterpret_cast<float4*>(square_avg),
      reinterpret_cast<float4*>(acc_delta), grad_norm, max_norm, lr, rho, eps, weight_decay, n4);
  return mmb::check_launch("adadelta_clip_kernel");
}
