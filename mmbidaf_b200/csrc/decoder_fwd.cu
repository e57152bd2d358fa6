// Fused multimodal-attention decoder step, forward (layers/attention.py:145-186).
//
// One decode step is three launches:
//   decoder_attn_kernel  (grid B)       two additive attentions with coverage (attention.py:147-156),
//                                       the 2-way modality soft-max (:161-166), att_cov and the coverage
//                                       update (:167, :177).  W1.enc / W3.enc are step invariant and arrive
//                                       pre-multiplied (proj_a / proj_i): the reference recomputes them
//                                       every step.
//   decoder_cell_kernel  (grid H/UPC)   the LSTM cell on [ctx, sent_embed] (:179-181), partitioned over
//                                       hidden units so each weight row is read once for the whole batch.
//   decoder_out_kernel   (grid B)       Linear(H -> M) + masked soft-max (:184) + first-max arg-max.
// The attention soft-maxes over the text axis are deliberately un-masked, as in the reference (quirk Q2).
#include "common.cuh"

namespace mmb {
namespace {

constexpr int ATT_THREADS = 512;
constexpr int CELL_THREADS = 256;
constexpr int OUT_THREADS = 256;
constexpr int UPC = 2;                 // hidden units per CTA in the cell kernel

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                      // protect `red` from a previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = lane < nw ? red[lane] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

struct AttnArgs {
  mmb_decoder_weights w;
  const float *proj_a, *proj_i, *enc_a, *enc_i, *h, *cov;
  float *ctx, *att_cov, *cov_out, *alpha, *beta, *ctx12;
  int B, Lt, H;
};

__global__ void __launch_bounds__(ATT_THREADS) decoder_attn_kernel(const AttnArgs a) {
  const int H = a.H, D = 2 * H, Lt = a.Lt;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = ATT_THREADS / 32;
  extern __shared__ __align__(16) float smem[];
  const bool vec4 = (D & 3) == 0;
  const int n_groups = vec4 ? ATT_THREADS / (D >> 2) : ATT_THREADS / D;
  float* part = smem;                      // [groups][2][D]  (first: 16-byte aligned for the float4 path)
  float* h_s = part + n_groups * 2 * D;    // [H]
  float* hw = h_s + H;                     // [4][D]  W2 h + b2 + bc1 | W4 h + b4 + bc2 | Wb2 h + bb2 | Wb4 h + bb4
  float* vec = hw + 4 * D;                 // [4][D]  v1 | wc1 | v2 | wc2
  float* ctx1 = vec + 4 * D;               // [D]
  float* ctx2 = ctx1 + D;                  // [D]
  float* red = ctx2 + D;                   // [32]
  float* e1 = red + 32;                    // [Lt]
  float* e2 = e1 + Lt;                     // [Lt]

  for (int i = tid; i < H; i += ATT_THREADS) h_s[i] = a.h[(size_t)b * H + i];
  for (int i = tid; i < D; i += ATT_THREADS) {
    vec[i] = a.w.v1[i];
    vec[D + i] = a.w.Wc1[i];
    vec[2 * D + i] = a.w.v2[i];
    vec[3 * D + i] = a.w.Wc2[i];
  }
  __syncthreads();
  // ---- the four h-side projections ------------------------------------------------------------------
#pragma unroll 4
  for (int r = warp; r < 4 * D; r += NW) {
    const int m = r / D, d = r - m * D;
    const float* W = m == 0 ? a.w.W2 : m == 1 ? a.w.W4 : m == 2 ? a.w.Wb2 : a.w.Wb4;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(W[(size_t)d * H + k], h_s[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float bias = m == 0 ? a.w.b2[d] + a.w.bc1[d] : m == 1 ? a.w.b4[d] + a.w.bc2[d] : m == 2 ? a.w.bb2[d] : a.w.bb4[d];
      hw[r] = acc + bias;
    }
  }
  __syncthreads();
  // ---- energies e_k[t] = v_k . tanh(proj_k[t] + W h + cov[t] wc_k) + v_k bias --------------------------
  const float v1b = a.w.v1b[0], v2b = a.w.v2b[0];
#pragma unroll 4
  for (int t = warp; t < Lt; t += NW) {
    const float* pa = a.proj_a + ((size_t)b * Lt + t) * D;
    const float* pi = a.proj_i + ((size_t)b * Lt + t) * D;
    const float c = a.cov[(size_t)b * Lt + t];
    float s1 = 0.f, s2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      s1 = fmaf(vec[d], tanh_fast((pa[d] + hw[d]) + c * vec[D + d]), s1);
      s2 = fmaf(vec[2 * D + d], tanh_fast((pi[d] + hw[D + d]) + c * vec[3 * D + d]), s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      e1[t] = s1 + v1b;
      e2[t] = s2 + v2b;
    }
  }
  __syncthreads();
  // ---- un-masked soft-max over the text axis (attention.py:148, :154) ---------------------------------
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int t = tid; t < Lt; t += ATT_THREADS) {
    m1 = fmaxf(m1, e1[t]);
    m2 = fmaxf(m2, e2[t]);
  }
  m1 = block_reduce(m1, red, true);
  m2 = block_reduce(m2, red, true);
  float l1 = 0.f, l2 = 0.f;
  for (int t = tid; t < Lt; t += ATT_THREADS) {
    const float p1 = expf(e1[t] - m1), p2 = expf(e2[t] - m2);
    e1[t] = p1;
    e2[t] = p2;
    l1 += p1;
    l2 += p2;
  }
  l1 = block_reduce(l1, red, false);
  l2 = block_reduce(l2, red, false);
  const float inv1 = 1.f / l1, inv2 = 1.f / l2;
  for (int t = tid; t < Lt; t += ATT_THREADS) {
    e1[t] *= inv1;
    e2[t] *= inv2;
  }
  __syncthreads();
  // ---- contexts c_k = sum_t alpha_k[t] enc_k[t] --------------------------------------------------------
  const float* ea = a.enc_a + (size_t)b * Lt * D;
  const float* ei = a.enc_i + (size_t)b * Lt * D;
  const int groups = n_groups;
  if (vec4) {
    const int dv4 = D >> 2;
    const int g = tid / dv4, c4 = tid - g * dv4;
    if (g < groups) {
      float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll 8
      for (int t = g; t < Lt; t += groups) {
        const float w1 = e1[t], w2 = e2[t];
        const float4 x1 = *reinterpret_cast<const float4*>(ea + (size_t)t * D + c4 * 4);
        const float4 x2 = *reinterpret_cast<const float4*>(ei + (size_t)t * D + c4 * 4);
        s1.x = fmaf(w1, x1.x, s1.x); s1.y = fmaf(w1, x1.y, s1.y); s1.z = fmaf(w1, x1.z, s1.z); s1.w = fmaf(w1, x1.w, s1.w);
        s2.x = fmaf(w2, x2.x, s2.x); s2.y = fmaf(w2, x2.y, s2.y); s2.z = fmaf(w2, x2.z, s2.z); s2.w = fmaf(w2, x2.w, s2.w);
      }
      *reinterpret_cast<float4*>(part + (g * 2 + 0) * D + c4 * 4) = s1;
      *reinterpret_cast<float4*>(part + (g * 2 + 1) * D + c4 * 4) = s2;
    }
  } else {
    const int g = tid / D, d = tid - g * D;
    if (g < groups) {
      float s1 = 0.f, s2 = 0.f;
      for (int t = g; t < Lt; t += groups) {
        s1 = fmaf(e1[t], ea[(size_t)t * D + d], s1);
        s2 = fmaf(e2[t], ei[(size_t)t * D + d], s2);
      }
      part[(g * 2 + 0) * D + d] = s1;
      part[(g * 2 + 1) * D + d] = s2;
    }
  }
  __syncthreads();
  for (int d = tid; d < D; d += ATT_THREADS) {
    float s1 = 0.f, s2 = 0.f;
    for (int g = 0; g < groups; ++g) {
      s1 += part[(g * 2 + 0) * D + d];
      s2 += part[(g * 2 + 1) * D + d];
    }
    ctx1[d] = s1;
    ctx2[d] = s2;
  }
  __syncthreads();
  // ---- modality attention beta (attention.py:161-164) ---------------------------------------------------
  float eb1 = 0.f, eb2 = 0.f;
#pragma unroll 4
  for (int r = warp; r < 2 * D; r += NW) {
    const int m = r / D, d = r - m * D;
    const float* W = m == 0 ? a.w.Wb1 : a.w.Wb3;
    const float* cx = m == 0 ? ctx1 : ctx2;
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(W[(size_t)d * D + k], cx[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (m == 0) eb1 = fmaf(a.w.vb1[d], tanh_fast((acc + a.w.bb1[d]) + hw[2 * D + d]), eb1);
      else eb2 = fmaf(a.w.vb2[d], tanh_fast((acc + a.w.bb3[d]) + hw[3 * D + d]), eb2);
    }
  }
  eb1 = block_reduce(eb1, red, false) + a.w.vb1b[0];
  eb2 = block_reduce(eb2, red, false) + a.w.vb2b[0];
  const float mb = fmaxf(eb1, eb2);
  const float x1 = expf(eb1 - mb), x2 = expf(eb2 - mb);
  const float beta1 = x1 / (x1 + x2), beta2 = x2 / (x1 + x2);
  for (int d = tid; d < D; d += ATT_THREADS) {
    a.ctx[(size_t)b * D + d] = ctx1[d] * beta1 + ctx2[d] * beta2;
    if (a.ctx12) {
      a.ctx12[((size_t)b * 2 + 0) * D + d] = ctx1[d];
      a.ctx12[((size_t)b * 2 + 1) * D + d] = ctx2[d];
    }
  }
  for (int t = tid; t < Lt; t += ATT_THREADS) {
    const float att = e1[t] * beta1 + e2[t] * beta2;                 // bmm([a1 a2], beta), attention.py:167
    a.att_cov[(size_t)b * Lt + t] = att;
    a.cov_out[(size_t)b * Lt + t] = a.cov[(size_t)b * Lt + t] + att;
    if (a.alpha) {
      a.alpha[((size_t)b * 2 + 0) * Lt + t] = e1[t];
      a.alpha[((size_t)b * 2 + 1) * Lt + t] = e2[t];
    }
  }
  if (a.beta && tid == 0) {
    a.beta[b * 2 + 0] = beta1;
    a.beta[b * 2 + 1] = beta2;
  }
}

struct CellArgs {
  const float *w_ih, *w_hh, *b_ih, *b_hh;       // (4H, D+E), (4H, H), (4H), (4H)
  const float *ctx, *sent, *h, *cell;           // (B,D) (B,E) (B,H) (B,H)
  float *h_out, *cell_out, *gates;              // (B,H) (B,H) (B,4H) activated gates (nullable)
  int B, H, E;
};

__global__ void __launch_bounds__(CELL_THREADS) decoder_cell_kernel(const CellArgs a) {
  const int H = a.H, D = 2 * H, E = a.E, K = D + E + H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = CELL_THREADS / 32;
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                              // [UPC*4][K]   rows (unit u, gate g): [W_ih | W_hh]
  float* bs = ws + UPC * 4 * K;                  // [UPC*4]
  const int j0 = blockIdx.x * UPC;
  for (int i = tid; i < UPC * 4 * K; i += CELL_THREADS) {
    const int row = i / K, k = i - row * K;
    const int u = row >> 2, g = row & 3, j = j0 + u;
    float v = 0.f;
    if (j < H) v = k < D + E ? a.w_ih[(size_t)(g * H + j) * (D + E) + k] : a.w_hh[(size_t)(g * H + j) * H + (k - D - E)];
    ws[i] = v;
  }
  for (int i = tid; i < UPC * 4; i += CELL_THREADS) {
    const int u = i >> 2, g = i & 3, j = j0 + u;
    bs[i] = j < H ? a.b_ih[g * H + j] + a.b_hh[g * H + j] : 0.f;
  }
  __syncthreads();
  for (int b = warp; b < a.B; b += NW) {
    float acc[UPC * 4];
#pragma unroll
    for (int r = 0; r < UPC * 4; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
      const float x = k < D ? a.ctx[(size_t)b * D + k] : k < D + E ? a.sent[(size_t)b * E + (k - D)]
                                                                 : a.h[(size_t)b * H + (k - D - E)];
#pragma unroll
      for (int r = 0; r < UPC * 4; ++r) acc[r] = fmaf(ws[r * K + k], x, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < UPC * 4; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < UPC && j0 + lane < H) {
      const int j = j0 + lane;
      float pre[4];
#pragma unroll
      for (int u = 0; u < UPC; ++u)
        if (u == lane) {
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[g] = acc[u * 4 + g] + bs[u * 4 + g];
        }
      const float gi = gate_act(pre[0], 1.f), gf = gate_act(pre[1], 1.f), gg = tanh_fast(pre[2]), go = gate_act(pre[3], 1.f);
      const float c = fmaf(gf, a.cell[(size_t)b * H + j], gi * gg);
      a.cell_out[(size_t)b * H + j] = c;
      a.h_out[(size_t)b * H + j] = go * tanh_fast(c);
      if (a.gates) {
        float* gs = a.gates + (size_t)b * 4 * H;
        gs[j] = gi; gs[H + j] = gf; gs[2 * H + j] = gg; gs[3 * H + j] = go;
      }
    }
  }
}

struct OutArgs {
  const float *w, *bias, *h;                    // (M,H) (M) (B,H)
  const uint8_t* mask;                          // (B,M)
  float* probs;                                 // (B,M)
  long long* argmax;                            // (B) nullable
  int B, H, M;
};

__global__ void __launch_bounds__(OUT_THREADS) decoder_out_kernel(const OutArgs a) {
  const int H = a.H, M = a.M;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = OUT_THREADS / 32;
  extern __shared__ __align__(16) float smem[];
  float* h_s = smem;                             // [H]
  float* red = h_s + H;                          // [32]
  float* lg = red + 32;                          // [M]
  for (int i = tid; i < H; i += OUT_THREADS) h_s[i] = a.h[(size_t)b * H + i];
  __syncthreads();
  const uint8_t* mk = a.mask + (size_t)b * M;
#pragma unroll 8
  for (int m = warp; m < M; m += NW) {
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(a.w[(size_t)m * H + k], h_s[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) lg[m] = mk[m] ? acc + a.bias[m] : kNegFill;      // attention.py:94
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int m = tid; m < M; m += OUT_THREADS) mx = fmaxf(mx, lg[m]);
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int m = tid; m < M; m += OUT_THREADS) {
    const float p = expf(lg[m] - mx);
    lg[m] = p;
    sum += p;
  }
  sum = block_reduce(sum, red, false);
  const float inv = 1.f / sum;
  float best = -INFINITY;
  int best_i = M;
  for (int m = tid; m < M; m += OUT_THREADS) {
    const float p = lg[m] * inv;
    a.probs[(size_t)b * M + m] = p;
    if (p > best) { best = p; best_i = m; }
  }
  if (a.argmax) {                                // first maximal index, as torch.max(dim) documents
    const float gbest = block_reduce(best, red, true);
    int cand = best == gbest ? best_i : M;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    int* redi = reinterpret_cast<int*>(red);
    __syncthreads();
    if (lane == 0) redi[warp] = cand;
    __syncthreads();
    if (tid == 0) {
      int r = M;
      for (int w = 0; w < NW; ++w) r = min(r, redi[w]);
      a.argmax[b] = r;
    }
  }
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_decoder_step_fwd(const mmb_decoder_weights* w, const float* proj_a, const float* proj_i,
                                    const float* enc_a, const float* enc_i, const float* sent_embed, const float* h,
                                    const float* cell, const float* coverage, const uint8_t* mask, float* probs,
                                    float* h_out, float* cell_out, float* att_cov, float* cov_out, long long* argmax,
                                    float* ctx, float* alpha, float* beta, float* gates, float* ctx12, int B, int Lt,
                                    int H, int E, int M, mmb_stream_t stream) {
  using namespace mmb;
  MMB_REQUIRE(w && proj_a && proj_i && enc_a && enc_i && sent_embed && h && cell && coverage && mask && probs && h_out &&
                  cell_out && att_cov && cov_out && ctx,
              MMB_ERR_INVALID, "mmb_decoder_step_fwd: null pointer");
  MMB_REQUIRE(B > 0 && Lt > 0 && H > 0 && E > 0 && M > 0, MMB_ERR_INVALID, "mmb_decoder_step_fwd: bad sizes");
  const int D = 2 * H;
  MMB_REQUIRE(D <= ATT_THREADS, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_fwd: hidden size %d > %d", H, ATT_THREADS / 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    AttnArgs a{*w, proj_a, proj_i, enc_a, enc_i, h, coverage, ctx, att_cov, cov_out, alpha, beta, ctx12, B, Lt, H};
    const int groups = (D % 4 == 0) ? ATT_THREADS / (D / 4) : ATT_THREADS / D;
    const size_t smem = sizeof(float) * ((size_t)H + 10 * D + 32 + 2 * (size_t)Lt + (size_t)groups * 2 * D);
    MMB_REQUIRE(smem <= 227 * 1024, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_fwd: Lt=%d needs %zu B of shared memory", Lt, smem);
    MMB_CUDA(cudaFuncSetAttribute(decoder_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decoder_attn_kernel<<<B, ATT_THREADS, smem, st>>>(a);
    if (int rc = check_launch("decoder_attn_kernel")) return rc;
  }
  {
    CellArgs a{w->lstm_w_ih, w->lstm_w_hh, w->lstm_b_ih, w->lstm_b_hh, ctx, sent_embed, h, cell, h_out, cell_out, gates, B, H, E};
    const size_t smem = sizeof(float) * ((size_t)UPC * 4 * (D + E + H) + UPC * 4);
    MMB_REQUIRE(smem <= 227 * 1024, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_fwd: E=%d too large", E);
    MMB_CUDA(cudaFuncSetAttribute(decoder_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decoder_cell_kernel<<<(H + UPC - 1) / UPC, CELL_THREADS, smem, st>>>(a);
    if (int rc = check_launch("decoder_cell_kernel")) return rc;
  }
  {
    OutArgs a{w->out_w, w->out_b, h_out, mask, probs, argmax, B, H, M};
    const size_t smem = sizeof(float) * ((size_t)H + 32 + M);
    MMB_REQUIRE(smem <= 227 * 1024, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_fwd: M=%d too large", M);
    MMB_CUDA(cudaFuncSetAttribute(decoder_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decoder_out_kernel<<<B, OUT_THREADS, smem, st>>>(a);
    if (int rc = check_launch("decoder_out_kernel")) return rc;
  }
  return MMB_OK;
}
