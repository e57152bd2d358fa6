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

// Gradient pack: up to PACK_SEGS per-parameter gradient tensors -> their places in the flat gradient buffer, ONE launch; the source
// pointers travel in the kernel parameters (captured by value in a CUDA graph, nothing to keep alive).  A null source zero-fills.
// (torch.cat of the ~110 pieces: 40 us for 12.8 MB on the tail of every step.)
constexpr int PACK_SEGS = 120, PACK_TILE = 4096;     // floats per block
struct PackSegs {
  const float* src[PACK_SEGS];
  long long dst_off[PACK_SEGS];
  int n[PACK_SEGS];                                  // elements to copy; the destination is zero-padded up to pad[i]
  int pad[PACK_SEGS];
};
static_assert(sizeof(PackSegs) <= 4000, "kernel parameter space");

__global__ void __launch_bounds__(256) pack_segments_kernel(const __grid_constant__ PackSegs s, float* __restrict__ dst) {
  const int seg = blockIdx.y;
  const int n = s.n[seg], pad = s.pad[seg];
  const int t0 = blockIdx.x * PACK_TILE;
  if (t0 >= pad) return;
  const float* src = s.src[seg];
  float* d = dst + s.dst_off[seg];
  const bool vec = src == nullptr || (reinterpret_cast<uintptr_t>(src) & 15) == 0;      // (dst_off and pad are multiples of 4)
  const int t1 = min(pad, t0 + PACK_TILE);
  if (vec) {
    for (int i = t0 + threadIdx.x * 4; i < t1; i += 256 * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src) {
        if (i + 4 <= n) v = __ldg(reinterpret_cast<const float4*>(src + i));
        else {
          float* e = reinterpret_cast<float*>(&v);
          for (int k = 0; k < 4; ++k) e[k] = i + k < n ? src[i + k] : 0.f;
        }
      }
      *reinterpret_cast<float4*>(d + i) = v;
    }
  } else {
    for (int i = t0 + threadIdx.x; i < t1; i += 256) d[i] = i < n ? src[i] : 0.f;
  }
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_pack_segments(const float* const* srcs, const long long* dst_offsets, const long long* sizes, int n_segs, float* dst,
                                 mmb_stream_t stream) {
  using namespace mmb;
  MMB_REQUIRE(srcs && dst_offsets && sizes && dst && n_segs > 0, MMB_ERR_INVALID, "mmb_pack_segments: bad arguments");
  MMB_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, MMB_ERR_UNSUPPORTED, "mmb_pack_segments: dst must be 16-byte aligned");
  for (int base = 0; base < n_segs; base += PACK_SEGS) {
    const int cnt = n_segs - base < PACK_SEGS ? n_segs - base : PACK_SEGS;
    PackSegs s{};
    int max_pad = 0;
    for (int i = 0; i < cnt; ++i) {
      const long long sz = sizes[base + i];
      MMB_REQUIRE(sz >= 0 && sz < (1ll << 31) - 8 && dst_offsets[base + i] % 4 == 0, MMB_ERR_INVALID,
                  "mmb_pack_segments: segment %d: size %lld, offset %lld (offsets must be multiples of 4)", base + i, sz,
                  dst_offsets[base + i]);
      s.src[i] = srcs[base + i];
      s.dst_off[i] = dst_offsets[base + i];
      s.n[i] = (int)sz;
      s.pad[i] = (int)((sz + 3) / 4 * 4);
      max_pad = s.pad[i] > max_pad ? s.pad[i] : max_pad;
    }
    if (max_pad == 0) continue;
    dim3 grid((unsigned)((max_pad + PACK_TILE - 1) / PACK_TILE), (unsigned)cnt);
    pack_segments_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(s, dst);
    if (int rc = check_launch("pack_segments_kernel")) return rc;
  }
  return MMB_OK;
}

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
