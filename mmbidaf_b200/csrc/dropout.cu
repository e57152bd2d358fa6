// Dropout of a plain tensor with the library's counter-based keep bits (common.cuh::dropout_keep): the reference's
// `F.dropout(x, drop_prob, training)` on the embedding inputs (layers/encoding.py:26) as ONE own launch -- no mask tensor is written,
// the backward pass (when the input needs a gradient) recomputes the bits from the same key.  The recurrent layers apply the same
// bits inside their kernels (csrc/bilstm.cu); the BiDAF kernels take the bits as a byte mask drawn by mmb_dropout_mask (bilstm.cu).
#include "common.cuh"

namespace mmb {
namespace {

// y[i] = keep(i) ? x[i] / keep_prob : 0.  128-bit accesses over the multiple-of-4 prefix, a scalar tail; x may equal y.
__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            const unsigned long long* __restrict__ rng_key, const float keep_prob,
                                                            const uint32_t n) {
  const unsigned long long key = rng_key[0];
  const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32), thresh = dropout_thresh(keep_prob);
  const float inv = 1.f / keep_prob;
  const uint32_t n4 = n >> 2, stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
    const uint32_t e = i << 2;
    v.x = dropout_keep(e, k0, k1, thresh) ? v.x * inv : 0.f;
    v.y = dropout_keep(e + 1, k0, k1, thresh) ? v.y * inv : 0.f;
    v.z = dropout_keep(e + 2, k0, k1, thresh) ? v.z * inv : 0.f;
    v.w = dropout_keep(e + 3, k0, k1, thresh) ? v.w * inv : 0.f;
    reinterpret_cast<float4*>(y)[i] = v;
  }
  for (uint32_t e = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
    y[e] = dropout_keep(e, k0, k1, thresh) ? x[e] * inv : 0.f;
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_dropout_apply(const float* x, float* y, const unsigned long long* rng_key, float keep_prob, long long n,
                                 mmb_stream_t stream) {
  MMB_REQUIRE(x && y && rng_key, MMB_ERR_INVALID, "mmb_dropout_apply: null pointer");
  MMB_REQUIRE(n > 0 && n < (1ll << 32) && keep_prob > 0.f && keep_prob <= 1.f, MMB_ERR_INVALID,
              "mmb_dropout_apply: n=%lld keep_prob=%g", n, keep_prob);
  MMB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, MMB_ERR_INVALID,
              "mmb_dropout_apply: x / y must be 16-byte aligned");
  const long long want = (n / 4 + 255) / 256;
  const unsigned blocks = (unsigned)(want < 1 ? 1 : (want < 148 * 8 ? want : 148 * 8));
  mmb::dropout_apply_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rng_key, keep_prob, (uint32_t)n);
  return mmb::check_launch("dropout_apply_kernel");
}
