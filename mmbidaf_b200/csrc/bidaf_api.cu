// C-ABI entry points for BiDAF attention (validation + precision dispatch).
#include "common.cuh"

namespace mmb {
int bidaf_fwd_f32(const float*, const float*, const uint8_t*, const uint8_t*, const float*, const float*, const float*,
                  const float*, const uint8_t*, const uint8_t*, float, float*, float*, float*, float*, float*, int, int, int,
                  int, cudaStream_t);
int bidaf_fwd_tc(const float*, const float*, const uint8_t*, const uint8_t*, const float*, const float*, const float*,
                 const float*, const uint8_t*, const uint8_t*, float, float*, float*, float*, float*, float*, void*, int, int,
                 int, int, cudaStream_t);
size_t bidaf_tc_workspace_bytes(int B, int Lc, int Lq, int dropout);
size_t bidaf_bwd_tc_workspace_bytes(int B, int Lc, int Lq);
size_t bidaf_bwd_f32_workspace_bytes(int B, int Lc, int Lq, int d);
int bidaf_bwd_f32(const float* grad_out, const float* text, const float* modality, const uint8_t* text_mask,
                  const uint8_t* modality_mask, const float* w_text, const float* w_modality, const float* w_cross,
                  const float* bias, const uint8_t* keep_text, const uint8_t* keep_modality, float keep_scale, const float* out,
                  const float* bm, const float* q2c, const float* lse_row, const float* lse_col, void* ws, float* d_text,
                  float* d_modality, float* d_w_text, float* d_w_modality, float* d_w_cross, float* d_bias, int B, int Lc, int Lq,
                  int d, cudaStream_t stream);
int bidaf_bwd_tc(const float* grad_out, const float* text, const float* modality, const float* w_text,
                 const float* w_modality, const float* w_cross, const float* bias, const uint8_t* keep_text,
                 const uint8_t* keep_modality, float keep_scale, const float* out, const float* bm, const float* q2c,
                 const float* lse_row, const float* lse_col, const void* fwd_workspace, void* workspace, float* d_text,
                 float* d_modality, float* d_w_text, float* d_w_modality, float* d_w_cross, float* d_bias, int B, int Lc,
                 int Lq, int d, cudaStream_t stream);
}

extern "C" size_t mmb_bidaf_workspace_bytes(int B, int Lc, int Lq, int d, int precision, int dropout) {
  (void)d;
  if (precision != MMB_PREC_BF16 || B <= 0 || Lc <= 0 || Lq <= 0) return 0;
  return mmb::bidaf_tc_workspace_bytes(B, Lc, Lq, dropout);
}

extern "C" int mmb_bidaf_fwd(const float* text, const float* modality, const uint8_t* text_mask,
                             const uint8_t* modality_mask, const float* w_text, const float* w_modality,
                             const float* w_cross, const float* bias, const uint8_t* keep_text,
                             const uint8_t* keep_modality, float keep_scale, float* out, float* q2c, float* bm,
                             float* lse_row, float* lse_col, void* workspace, int B, int Lc, int Lq, int d, int precision,
                             mmb_stream_t stream) {
  MMB_REQUIRE(text && modality && text_mask && modality_mask && w_text && w_modality && w_cross && bias && out, MMB_ERR_INVALID,
              "mmb_bidaf_fwd: null pointer");
  // q2c, lse_row and lse_col (what the backward pass wants saved) may be NULL on the bf16 tier: inference writes `out` only
  MMB_REQUIRE(precision == MMB_PREC_BF16 || (q2c && lse_row && lse_col), MMB_ERR_INVALID,
              "mmb_bidaf_fwd: q2c / lse_row / lse_col are required on the fp32 tier");
  MMB_REQUIRE(B > 0 && Lc > 0 && Lq > 0 && d > 0, MMB_ERR_INVALID, "mmb_bidaf_fwd: B=%d Lc=%d Lq=%d d=%d", B, Lc, Lq, d);
  MMB_REQUIRE(d % 4 == 0 && d <= 256, MMB_ERR_UNSUPPORTED, "mmb_bidaf_fwd: d=%d (need d %% 4 == 0, d <= 256)", d);
  MMB_REQUIRE(B <= 65535, MMB_ERR_UNSUPPORTED, "mmb_bidaf_fwd: B=%d > 65535", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == MMB_PREC_FP32)
    return mmb::bidaf_fwd_f32(text, modality, text_mask, modality_mask, w_text, w_modality, w_cross, bias, keep_text,
                              keep_modality, keep_scale, out, q2c, bm, lse_row, lse_col, B, Lc, Lq, d, st);
  if (precision == MMB_PREC_BF16)
    return mmb::bidaf_fwd_tc(text, modality, text_mask, modality_mask, w_text, w_modality, w_cross, bias, keep_text,
                             keep_modality, keep_scale, out, q2c, bm, lse_row, lse_col, workspace, B, Lc, Lq, d, st);
  mmb::set_error("mmb_bidaf_fwd: precision %d not available", precision);
  return MMB_ERR_UNSUPPORTED;
}

extern "C" size_t mmb_bidaf_bwd_workspace_bytes(int B, int Lc, int Lq, int d, int precision) {
  if (B <= 0 || Lc <= 0 || Lq <= 0 || d <= 0) return 0;
  if (precision == MMB_PREC_FP32) return mmb::bidaf_bwd_f32_workspace_bytes(B, Lc, Lq, d);
  if (precision != MMB_PREC_BF16) return 0;
  return mmb::bidaf_bwd_tc_workspace_bytes(B, Lc, Lq);
}

extern "C" int mmb_bidaf_bwd(const float* grad_out, const float* text, const float* modality, const uint8_t* text_mask,
                             const uint8_t* modality_mask, const float* w_text,
                             const float* w_modality, const float* w_cross, const float* bias, const uint8_t* keep_text,
                             const uint8_t* keep_modality, float keep_scale, const float* out, const float* bm,
                             const float* q2c, const float* lse_row, const float* lse_col, const void* fwd_workspace,
                             void* workspace, float* d_text, float* d_modality, float* d_w_text, float* d_w_modality,
                             float* d_w_cross, float* d_bias, int B, int Lc, int Lq, int d, int precision,
                             mmb_stream_t stream) {
  MMB_REQUIRE(grad_out && text && modality && w_text && w_modality && w_cross && bias && out && bm && q2c && lse_row &&
                  lse_col && d_text && d_modality && d_w_text && d_w_modality && d_w_cross && d_bias,
              MMB_ERR_INVALID, "mmb_bidaf_bwd: null pointer");
  MMB_REQUIRE(B > 0 && Lc > 0 && Lq > 0 && d > 0, MMB_ERR_INVALID, "mmb_bidaf_bwd: B=%d Lc=%d Lq=%d d=%d", B, Lc, Lq, d);
  MMB_REQUIRE(B <= 65535, MMB_ERR_UNSUPPORTED, "mmb_bidaf_bwd: B=%d > 65535", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == MMB_PREC_FP32)
    return mmb::bidaf_bwd_f32(grad_out, text, modality, text_mask, modality_mask, w_text, w_modality, w_cross, bias, keep_text,
                              keep_modality, keep_scale, out, bm, q2c, lse_row, lse_col, workspace, d_text, d_modality, d_w_text,
                              d_w_modality, d_w_cross, d_bias, B, Lc, Lq, d, st);
  if (precision == MMB_PREC_BF16)
    return mmb::bidaf_bwd_tc(grad_out, text, modality, w_text, w_modality, w_cross, bias, keep_text, keep_modality,
                             keep_scale, out, bm, q2c, lse_row, lse_col, fwd_workspace, workspace, d_text, d_modality,
                             d_w_text, d_w_modality, d_w_cross, d_bias, B, Lc, Lq, d, st);
  mmb::set_error("mmb_bidaf_bwd: precision %d not available", precision);
  return MMB_ERR_UNSUPPORTED;
}
