// Backward of one multimodal-attention decoder step (layers/attention.py:145-186), two launches:
//
//   decoder_tail_bwd_kernel (grid B)  masked soft-max + Linear(H->M) + LSTM cell: produces d_logits, the
//                                     gate pre-activation gradients, d_cell, d_ctx and the LSTM part of d_h.
//   decoder_attn_bwd_kernel (grid B)  modality soft-max, W_beta paths, both additive attentions with
//                                     coverage: produces d_cov, adds the attention part of d_h, accumulates
//                                     d_proj_* in place and emits the rows from which the caller forms all
//                                     weight gradients with ONE set of GEMMs per sequence (not per step).
//
// Gradients of the un-masked attention soft-maxes, of the 2-way soft-max, of tanh and of the LSTM cell are
// written out explicitly; the derivation was checked against the reference's autograd (tests/golden/
// decoder_small.pt, model_small.pt).
#include "common.cuh"

namespace mmb {
namespace {

constexpr int TAIL_THREADS = 256;
constexpr int ATTB_THREADS = 512;

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = lane < nw ? red[lane] : 0.f;
  return warp_sum(r);
}

struct TailArgs {
  const float *out_w, *w_ih, *w_hh;              // (M,H) (4H,D+E) (4H,H)
  const float *probs, *gates, *cell_in, *cell_out;
  const float *d_probs, *d_h_out, *d_cell_out;   // d_probs / d_cell_out may be null (zero)
  float *d_logits, *d_gates, *d_cell, *d_ctx, *d_h;
  int B, H, E, M;
};

__global__ void __launch_bounds__(TAIL_THREADS) decoder_tail_bwd_kernel(const TailArgs a) {
  const int H = a.H, D = 2 * H, E = a.E, M = a.M;
  const int b = blockIdx.x, tid = threadIdx.x;
  extern __shared__ __align__(16) float smem[];
  float* red = smem;                 // [32]
  float* dh = red + 32;              // [H]
  float* da = dh + H;                // [4H]
  float* dlog = da + 4 * H;          // [M]
  // ---- masked soft-max backward: dlogit = p (dp - sum p dp); masked entries have p = 0 ------------------
  float dot = 0.f;
  if (a.d_probs)
    for (int m = tid; m < M; m += TAIL_THREADS) dot += a.probs[(size_t)b * M + m] * a.d_probs[(size_t)b * M + m];
  dot = block_sum(dot, red);
  for (int m = tid; m < M; m += TAIL_THREADS) {
    const float p = a.probs[(size_t)b * M + m];
    const float g = a.d_probs ? p * (a.d_probs[(size_t)b * M + m] - dot) : 0.f;
    dlog[m] = g;
    a.d_logits[(size_t)b * M + m] = g;
  }
  __syncthreads();
  // ---- d h' = upstream + W_out^T dlogit ---------------------------------------------------------------
  for (int k = tid; k < H; k += TAIL_THREADS) {
    float acc = a.d_h_out ? a.d_h_out[(size_t)b * H + k] : 0.f;
#pragma unroll 16
    for (int m = 0; m < M; ++m) acc = fmaf(dlog[m], a.out_w[(size_t)m * H + k], acc);
    dh[k] = acc;
  }
  __syncthreads();
  // ---- LSTM cell backward ---------------------------------------------------------------------------------
  for (int k = tid; k < H; k += TAIL_THREADS) {
    const float* g = a.gates + (size_t)b * 4 * H;
    const float gi = g[k], gf = g[H + k], gg = g[2 * H + k], go = g[3 * H + k];
    const float tc = tanh_fast(a.cell_out[(size_t)b * H + k]);
    const float dc = fmaf(dh[k] * go, 1.f - tc * tc, a.d_cell_out ? a.d_cell_out[(size_t)b * H + k] : 0.f);
    const float ai = dc * gg * gi * (1.f - gi);
    const float af = dc * a.cell_in[(size_t)b * H + k] * gf * (1.f - gf);
    const float ag = dc * gi * (1.f - gg * gg);
    const float ao = dh[k] * tc * go * (1.f - go);
    da[k] = ai; da[H + k] = af; da[2 * H + k] = ag; da[3 * H + k] = ao;
    float* o = a.d_gates + (size_t)b * 4 * H;
    o[k] = ai; o[H + k] = af; o[2 * H + k] = ag; o[3 * H + k] = ao;
    a.d_cell[(size_t)b * H + k] = dc * gf;
  }
  __syncthreads();
  // ---- d ctx = (W_ih^T da)[:D]   (sent_embed needs no gradient),  d h (LSTM part) = W_hh^T da ----------------
  for (int d = tid; d < D + H; d += TAIL_THREADS) {
    float acc = 0.f;
    if (d < D) {
#pragma unroll 16
      for (int r = 0; r < 4 * H; ++r) acc = fmaf(da[r], a.w_ih[(size_t)r * (D + E) + d], acc);
      a.d_ctx[(size_t)b * D + d] = acc;
    } else {
      const int k = d - D;
#pragma unroll 16
      for (int r = 0; r < 4 * H; ++r) acc = fmaf(da[r], a.w_hh[(size_t)r * H + k], acc);
      a.d_h[(size_t)b * H + k] = acc;
    }
  }
}

struct AttnBwdArgs {
  mmb_decoder_weights w;
  const float *proj_a, *proj_i, *enc_a, *enc_i, *h, *cov;
  const float *alpha, *beta, *ctx12;             // saved: (B,2,Lt) (B,2) (B,2,D)
  const float *d_ctx, *d_att_cov, *d_cov_out;    // (B,D) (B,Lt)|null (B,Lt)|null
  float *d_h, *d_cov, *d_proj_a, *d_proj_i, *d_ctx12, *d_pre, *vec_acc, *scal_acc;
  int B, Lt, H;
};

__global__ void __launch_bounds__(ATTB_THREADS) decoder_attn_bwd_kernel(const AttnBwdArgs a) {
  const int H = a.H, D = 2 * H, Lt = a.Lt;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = ATTB_THREADS / 32;
  extern __shared__ __align__(16) float smem[];
  float* red = smem;                     // [32]
  float* h_s = red + 32;                 // [H]
  float* hw = h_s + H;                   // [4][D]   as in the forward kernel
  float* vec = hw + 4 * D;               // [4][D]   v1 | wc1 | v2 | wc2
  float* ctx = vec + 4 * D;              // [2][D]   c1 | c2
  float* dctx = ctx + 2 * D;             // [2][D]   d c1 | d c2
  float* dpre = dctx + 2 * D;            // [4][D]   d(W2 h) | d(W4 h) | d pre_beta1 | d pre_beta3
  float* al = dpre + 4 * D;              // [2][Lt]  alpha, then d e
  float* datt = al + 2 * Lt;             // [Lt]     d att_cov + d cov_out
  float* part = datt + Lt;               // [NW][3][D] per-warp column partials

  const float beta1 = a.beta[b * 2 + 0], beta2 = a.beta[b * 2 + 1];
  for (int i = tid; i < H; i += ATTB_THREADS) h_s[i] = a.h[(size_t)b * H + i];
  for (int i = tid; i < D; i += ATTB_THREADS) {
    vec[i] = a.w.v1[i];
    vec[D + i] = a.w.Wc1[i];
    vec[2 * D + i] = a.w.v2[i];
    vec[3 * D + i] = a.w.Wc2[i];
    ctx[i] = a.ctx12[((size_t)b * 2 + 0) * D + i];
    ctx[D + i] = a.ctx12[((size_t)b * 2 + 1) * D + i];
  }
  for (int t = tid; t < Lt; t += ATTB_THREADS) {
    al[t] = a.alpha[((size_t)b * 2 + 0) * Lt + t];
    al[Lt + t] = a.alpha[((size_t)b * 2 + 1) * Lt + t];
    datt[t] = (a.d_att_cov ? a.d_att_cov[(size_t)b * Lt + t] : 0.f) + (a.d_cov_out ? a.d_cov_out[(size_t)b * Lt + t] : 0.f);
  }
  __syncthreads();
  // ---- recompute the h-side projections (as forward) ---------------------------------------------------------
#pragma unroll 4
  for (int r = warp; r < 4 * D; r += NW) {
    const int m = r / D, d = r - m * D;
    const float* W = m == 0 ? a.w.W2 : m == 1 ? a.w.W4 : m == 2 ? a.w.Wb2 : a.w.Wb4;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(W[(size_t)d * H + k], h_s[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float bias = m == 0 ? a.w.b2[d] + a.w.bc1[d] : m == 1 ? a.w.b4[d] + a.w.bc2[d] : m == 2 ? a.w.bb2[d] : a.w.bb4[d];
      hw[r] = acc + bias;
    }
  }
  // ---- d beta and the 2-way soft-max ------------------------------------------------------------------------------
  float db1 = 0.f, db2 = 0.f;
  for (int d = tid; d < D; d += ATTB_THREADS) {
    const float g = a.d_ctx[(size_t)b * D + d];
    db1 = fmaf(ctx[d], g, db1);
    db2 = fmaf(ctx[D + d], g, db2);
  }
  for (int t = tid; t < Lt; t += ATTB_THREADS) {
    db1 = fmaf(al[t], datt[t], db1);
    db2 = fmaf(al[Lt + t], datt[t], db2);
  }
  db1 = block_sum(db1, red);
  db2 = block_sum(db2, red);
  const float mix = beta1 * db1 + beta2 * db2;
  const float deb1 = beta1 * (db1 - mix), deb2 = beta2 * (db2 - mix);
  // ---- W_beta paths: pre_k = W_beta_{1,3} c_k + bias + (W_beta_{2,4} h + bias) -------------------------------------
#pragma unroll 4
  for (int r = warp; r < 2 * D; r += NW) {
    const int m = r / D, d = r - m * D;
    const float* W = m == 0 ? a.w.Wb1 : a.w.Wb3;
    const float* cx = ctx + m * D;
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(W[(size_t)d * D + k], cx[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float tb = tanh_fast((acc + (m == 0 ? a.w.bb1[d] : a.w.bb3[d])) + hw[(2 + m) * D + d]);
      const float de = m == 0 ? deb1 : deb2;
      const float vb = m == 0 ? a.w.vb1[d] : a.w.vb2[d];
      dpre[(2 + m) * D + d] = de * vb * (1.f - tb * tb);
      a.vec_acc[((size_t)b * 6 + 4 + m) * D + d] += de * tb;           // d v_beta weight
    }
  }
  if (tid == 0) {
    a.scal_acc[b * 4 + 2] += deb1;
    a.scal_acc[b * 4 + 3] += deb2;
  }
  __syncthreads();
  // d c_k = beta_k d ctx + W_beta^T d pre_k
  for (int i = tid; i < 2 * D; i += ATTB_THREADS) {
    const int m = i / D, d = i - m * D;
    const float* W = m == 0 ? a.w.Wb1 : a.w.Wb3;
    const float* dp = dpre + (2 + m) * D;
    float acc = (m == 0 ? beta1 : beta2) * a.d_ctx[(size_t)b * D + d];
#pragma unroll 16
    for (int r = 0; r < D; ++r) acc = fmaf(W[(size_t)r * D + d], dp[r], acc);
    dctx[i] = acc;
    a.d_ctx12[((size_t)b * 2 + m) * D + d] = acc;
  }
  __syncthreads();
  // ---- sweep 1 over enc: d alpha_k[t] = beta_k datt[t] + d c_k . enc_k[t]; soft-max backward over t ------------------
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
  for (int t = warp; t < Lt; t += NW) {
    const float* ea = a.enc_a + ((size_t)b * Lt + t) * D;
    const float* ei = a.enc_i + ((size_t)b * Lt + t) * D;
    float d1 = 0.f, d2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      d1 = fmaf(dctx[d], ea[d], d1);
      d2 = fmaf(dctx[D + d], ei[d], d2);
    }
    d1 = warp_sum(d1) + beta1 * datt[t];
    d2 = warp_sum(d2) + beta2 * datt[t];
    if (lane == 0) {
      const float a1 = al[t], a2 = al[Lt + t];
      s1 = fmaf(a1, d1, s1);
      s2 = fmaf(a2, d2, s2);
      al[t] = d1;                       // temporarily d alpha; turned into d e below
      al[Lt + t] = d2;
    }
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  float se1 = 0.f, se2 = 0.f;
  for (int t = tid; t < Lt; t += ATTB_THREADS) {
    const float e1 = a.alpha[((size_t)b * 2 + 0) * Lt + t] * (al[t] - s1);
    const float e2 = a.alpha[((size_t)b * 2 + 1) * Lt + t] * (al[Lt + t] - s2);
    al[t] = e1;
    al[Lt + t] = e2;
    se1 += e1;
    se2 += e2;
  }
  se1 = block_sum(se1, red);
  se2 = block_sum(se2, red);
  if (tid == 0) {
    a.scal_acc[b * 4 + 0] += se1;       // d v1 bias (identically 0 up to rounding)
    a.scal_acc[b * 4 + 1] += se2;
  }
  __syncthreads();
  // ---- sweep 2 over proj: z = proj + (W h + b) + cov wc;  dz = de v (1 - tanh^2 z) -----------------------------------
  for (int m = 0; m < 2; ++m) {
    const float* proj = (m == 0 ? a.proj_a : a.proj_i) + (size_t)b * Lt * D;
    float* dproj = (m == 0 ? a.d_proj_a : a.d_proj_i) + (size_t)b * Lt * D;
    const float* vv = vec + (2 * m) * D;
    const float* wc = vec + (2 * m + 1) * D;
    const float* hwm = hw + m * D;
    const float* de = al + m * Lt;
    // lane owns columns d = lane + 32 j; per-warp partial column sums live in registers across its rows
    constexpr int MAXJ = 8;              // D <= 256
    float c_dz[MAXJ], c_cov[MAXJ], c_v[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) c_dz[j] = c_cov[j] = c_v[j] = 0.f;
#pragma unroll 2
    for (int t = warp; t < Lt; t += NW) {
      const float cv = a.cov[(size_t)b * Lt + t];
      const float det = de[t];
      float row = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int d = lane + 32 * j;
        if (d < D) {
          const float tz = tanh_fast((proj[(size_t)t * D + d] + hwm[d]) + cv * wc[d]);
          const float dz = det * vv[d] * (1.f - tz * tz);
          dproj[(size_t)t * D + d] += dz;
          c_dz[j] += dz;
          c_cov[j] = fmaf(dz, cv, c_cov[j]);
          c_v[j] = fmaf(det, tz, c_v[j]);
          row = fmaf(dz, wc[d], row);
        }
      }
      row = warp_sum(row);
      if (lane == 0) {
        if (m == 0) datt[t] = (a.d_cov_out ? a.d_cov_out[(size_t)b * Lt + t] : 0.f) + row;   // becomes d cov
        else datt[t] += row;
      }
    }
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int d = lane + 32 * j;
      if (d < D) {
        part[(warp * 3 + 0) * D + d] = c_dz[j];
        part[(warp * 3 + 1) * D + d] = c_cov[j];
        part[(warp * 3 + 2) * D + d] = c_v[j];
      }
    }
    __syncthreads();
    for (int d = tid; d < D; d += ATTB_THREADS) {
      float x = 0.f, y = 0.f, z = 0.f;
      for (int w = 0; w < NW; ++w) {
        x += part[(w * 3 + 0) * D + d];
        y += part[(w * 3 + 1) * D + d];
        z += part[(w * 3 + 2) * D + d];
      }
      dpre[m * D + d] = x;                                             // d (W h + b)  -> dW2/dW4, biases
      a.vec_acc[((size_t)b * 6 + m) * D + d] += y;                      // d Wc weight
      a.vec_acc[((size_t)b * 6 + 2 + m) * D + d] += z;                  // d v weight
    }
    __syncthreads();
  }
  for (int t = tid; t < Lt; t += ATTB_THREADS) a.d_cov[(size_t)b * Lt + t] = datt[t];
  for (int i = tid; i < 4 * D; i += ATTB_THREADS) a.d_pre[(size_t)b * 4 * D + i] = dpre[i];
  // ---- attention part of d h: W2^T d0 + W4^T d1 + W_beta_2^T d2 + W_beta_4^T d3 ---------------------------------------
  for (int k = tid; k < H; k += ATTB_THREADS) {
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      acc = fmaf(a.w.W2[(size_t)d * H + k], dpre[d], acc);
      acc = fmaf(a.w.W4[(size_t)d * H + k], dpre[D + d], acc);
      acc = fmaf(a.w.Wb2[(size_t)d * H + k], dpre[2 * D + d], acc);
      acc = fmaf(a.w.Wb4[(size_t)d * H + k], dpre[3 * D + d], acc);
    }
    a.d_h[(size_t)b * H + k] += acc;
  }
}

}  // namespace
}  // namespace mmb

extern "C" int mmb_decoder_step_bwd(const mmb_decoder_weights* w, const float* proj_a, const float* proj_i,
                                    const float* enc_a, const float* enc_i, const float* h, const float* cell,
                                    const float* coverage, const float* probs, const float* h_out,
                                    const float* cell_out, const float* gates, const float* alpha, const float* beta,
                                    const float* ctx12, const float* d_probs, const float* d_h_out,
                                    const float* d_cell_out, const float* d_att_cov, const float* d_cov_out, float* d_h,
                                    float* d_cell, float* d_cov, float* d_proj_a, float* d_proj_i, float* d_logits,
                                    float* d_gates, float* d_ctx12, float* d_pre, float* vec_acc, float* scal_acc,
                                    float* d_ctx, int B, int Lt, int H, int E, int M, mmb_stream_t stream) {
  using namespace mmb;
  (void)h_out;
  MMB_REQUIRE(w && proj_a && proj_i && enc_a && enc_i && h && cell && coverage && probs && cell_out && gates && alpha &&
                  beta && ctx12 && d_h && d_cell && d_cov && d_proj_a && d_proj_i && d_logits && d_gates && d_ctx12 &&
                  d_pre && vec_acc && scal_acc && d_ctx,
              MMB_ERR_INVALID, "mmb_decoder_step_bwd: null pointer");
  MMB_REQUIRE(B > 0 && Lt > 0 && H > 0 && E > 0 && M > 0, MMB_ERR_INVALID, "mmb_decoder_step_bwd: bad sizes");
  const int D = 2 * H;
  MMB_REQUIRE(D <= 256, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_bwd: hidden size %d > 128", H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    TailArgs a{w->out_w, w->lstm_w_ih, w->lstm_w_hh, probs, gates, cell, cell_out, d_probs, d_h_out, d_cell_out,
               d_logits, d_gates, d_cell, d_ctx, d_h, B, H, E, M};
    const size_t smem = sizeof(float) * (32 + (size_t)H + 4 * H + M);
    MMB_REQUIRE(smem <= 227 * 1024, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_bwd: M=%d too large", M);
    MMB_CUDA(cudaFuncSetAttribute(decoder_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decoder_tail_bwd_kernel<<<B, TAIL_THREADS, smem, st>>>(a);
    if (int rc = check_launch("decoder_tail_bwd_kernel")) return rc;
  }
  {
    AttnBwdArgs a{*w, proj_a, proj_i, enc_a, enc_i, h, coverage, alpha, beta, ctx12, d_ctx, d_att_cov, d_cov_out,
                  d_h, d_cov, d_proj_a, d_proj_i, d_ctx12, d_pre, vec_acc, scal_acc, B, Lt, H};
    const size_t smem = sizeof(float) * (32 + (size_t)H + 16 * D + 3 * (size_t)Lt + (size_t)(ATTB_THREADS / 32) * 3 * D);
    MMB_REQUIRE(smem <= 227 * 1024, MMB_ERR_UNSUPPORTED, "mmb_decoder_step_bwd: Lt=%d needs %zu B of shared memory", Lt, smem);
    MMB_CUDA(cudaFuncSetAttribute(decoder_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decoder_attn_bwd_kernel<<<B, ATTB_THREADS, smem, st>>>(a);
    if (int rc = check_launch("decoder_attn_bwd_kernel")) return rc;
  }
  return MMB_OK;
}
