/*
 * mmbidaf_b200 -- C ABI of the B200 (sm_100a) kernels behind the MMBiDAF hot path.
 *
 * The reference (amankhullar/MMBiDAF) is pure Python: it has no FFI.  The boundary a
 * maintainer would bind is therefore the set of torch calls made inside
 *   layers/attention.py::BiDAFAttention.forward            (attention.py:37-75)
 *   layers/attention.py::masked_softmax                    (attention.py:78-98)
 *   layers/attention.py::MultimodalAttentionDecoder.forward (attention.py:145-186)
 *   layers/encoding.py::RNNEncoder.forward                 (encoding.py:83-108)
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows
 * the ctypes stub that binds them from the reference's own layers/*.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous, row-major memory unless it says
 *     "host"; floats are fp32; masks / keep-masks are one byte per element (0 or 1);
 *   - nothing here allocates, synchronises or keeps global mutable state: the caller
 *     owns outputs and workspaces and passes the CUDA stream to launch on;
 *   - every function returns 0 on success, a non-zero mmb_status otherwise, and leaves a
 *     message readable through mmb_last_error() (thread local);
 *   - unsupported shapes are an error (MMB_ERR_UNSUPPORTED); there is no CPU fallback.
 */
#ifndef MMBIDAF_B200_H_
#define MMBIDAF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MMB_API __attribute__((visibility("default")))
#else
#define MMB_API
#endif

typedef void* mmb_stream_t;          /* a cudaStream_t */

enum mmb_status {
  MMB_OK = 0,
  MMB_ERR_INVALID = 1,               /* null pointer / non-positive size */
  MMB_ERR_UNSUPPORTED = 2,           /* shape or precision this build has no kernel for */
  MMB_ERR_CUDA = 3                   /* a CUDA runtime call failed (message has the cudaError) */
};

enum mmb_precision {
  MMB_PREC_FP32 = 0,                 /* fp32 FFMA contractions: rel <= 1e-5 tier */
  MMB_PREC_BF16 = 1                  /* tcgen05 bf16 contractions, fp32 accumulate: rel <= 2e-2 tier */
};

MMB_API int mmb_version(void);
MMB_API const char* mmb_last_error(void);
/* 1 if the current device is sm_100 (B200); the python loader refuses anything else. */
MMB_API int mmb_device_supported(void);

/* --------------------------------------------------------------------------------------
 * BiDAF attention, forward.  Replaces attention.py:37-54 + :56-75 + two masked_softmax
 * calls (:43-44).  S is never written to memory.
 *   text (B,Lc,d)  modality (B,Lq,d)  text_mask (B,Lc)  modality_mask (B,Lq)
 *   w_text (d) = text_weight, w_modality (d) = modality_weight, w_cross (d) =
 *   text_modality_weight, bias (1)                              (attention.py:30-35)
 *   keep_text / keep_modality: optional (NULL in eval) dropout keep-masks (B,L,d) for the
 *   inputs of the similarity only (attention.py:66-67); keep_scale = 1/(1-p).
 * Outputs
 *   out (B,Lc,4d) = [c, a, c*a, c*b]                                    (attention.py:52)
 *   q2c (B,Lq,d) = s2^T c, lse_row (B,Lc), lse_col (B,Lq): log-sum-exp of the row / column
 *   soft-max -- saved for the backward pass.  MMB_PREC_FP32: d % 4 == 0, d <= 256 (workspace may be NULL).
 *   MMB_PREC_BF16 (tcgen05 + TMEM + TMA): d % 8 == 0, d <= 200, workspace of mmb_bidaf_workspace_bytes().
 */
MMB_API int mmb_bidaf_fwd(const float* text, const float* modality, const uint8_t* text_mask, const uint8_t* modality_mask,
                  const float* w_text, const float* w_modality, const float* w_cross, const float* bias,
                  const uint8_t* keep_text, const uint8_t* keep_modality, float keep_scale,
                  float* out, float* q2c, float* lse_row, float* lse_col, void* workspace,
                  int B, int Lc, int Lq, int d, int precision, mmb_stream_t stream);

/* Bytes of scratch `workspace` mmb_bidaf_fwd needs (0 for MMB_PREC_FP32).  The bf16 tier keeps bf16 copies of
 * its operands there in tensor-core order; dropout != 0 when keep_modality will be non-NULL. */
MMB_API size_t mmb_bidaf_workspace_bytes(int B, int Lc, int Lq, int d, int precision, int dropout);

/* --------------------------------------------------------------------------------------
 * Length-aware LSTM recurrence of one layer, 1 or 2 directions.  Replaces the nn.LSTM call of
 * encoding.py:96 together with the sort / pack / pad / unsort gathers of encoding.py:91-101.
 *   gates (B,L,ndir,4H): on entry x W_ih^T + b_ih + b_hh (gate order i,f,g,o; a plain GEMM done
 *     by the caller); with save != 0 it holds the ACTIVATED gates on exit (for the backward pass)
 *   w_hh (ndir,4H,H)   lengths (B) int32   order (B) int32 permutation, longest first, or NULL
 *   out (B,L,ndir*H): hidden states, exact zeros past each length (pad_packed_sequence)
 *   h_n, c_n (B,ndir,H): state after each sample's last valid step, in BATCH order
 *   cell (B,L,ndir,H): cell states, written when save != 0 (may be NULL otherwise).  H <= 104.
 */
MMB_API int mmb_bilstm_fwd(float* gates, const float* w_hh, const int32_t* lengths, const int32_t* order, float* out,
                           float* h_n, float* c_n, float* cell, int B, int L, int H, int ndir, int save,
                           mmb_stream_t stream);

/* Backward through time of the same layer.  `gates` (activated gates from the forward pass) is
 * overwritten with d(loss)/d(pre-activation) (zeros past each length); dout (B,L,ndir*H);
 * dh_n / dc_n (B,ndir,H) may be NULL.  dW_ih, dx, db and dW_hh are GEMMs over `gates` for the caller.
 */
MMB_API int mmb_bilstm_bwd(float* gates, const float* cell, const float* w_hh, const int32_t* lengths,
                           const int32_t* order, const float* dout, const float* dh_n, const float* dc_n, int B, int L,
                           int H, int ndir, mmb_stream_t stream);

/* --------------------------------------------------------------------------------------
 * Multimodal attention decoder, one step.  Replaces attention.py:145-186.
 * Device pointers to the module's parameters, named as in attention.py:119-142 (weights are
 * nn.Linear layout (out,in); *b are the biases).
 */
typedef struct mmb_decoder_weights {
  const float *W2, *b2, *Wc1, *bc1, *v1, *v1b;           /* text-audio additive attention (W1 is pre-applied) */
  const float *W4, *b4, *Wc2, *bc2, *v2, *v2b;           /* text-image additive attention (W3 is pre-applied) */
  const float *Wb1, *bb1, *Wb2, *bb2, *Wb3, *bb3, *Wb4, *bb4, *vb1, *vb1b, *vb2, *vb2b;   /* W_beta_1..4, v_beta_1..2 */
  const float *lstm_w_ih, *lstm_w_hh, *lstm_b_ih, *lstm_b_hh;   /* lstm.weight_ih_l0 (4H, 2H+E) ... */
  const float *out_w, *out_b;                            /* out: Linear(H -> M) */
} mmb_decoder_weights;

/*   proj_a = W1(enc_a) + b1, proj_i = W3(enc_i) + b3 (B,Lt,2H): step-invariant GEMMs done once by the caller
 *   enc_a, enc_i (B,Lt,2H)   sent_embed (B,E)   h, cell (B,H)   coverage (B,Lt)   mask (B,M)
 * Outputs: probs (B,M) = masked_softmax(out(h')) (attention.py:184); h_out, cell_out (B,H);
 *   att_cov, cov_out (B,Lt) (attention.py:167,177); argmax (B) int64 first-max index of probs (nullable);
 *   ctx (B,2H) the attended context (also a workspace); alpha (B,2,Lt), beta (B,2), gates (B,4H): nullable,
 *   saved for a backward pass.  `w` is a HOST pointer to the struct.
 */
MMB_API int mmb_decoder_step_fwd(const mmb_decoder_weights* w, const float* proj_a, const float* proj_i,
                                 const float* enc_a, const float* enc_i, const float* sent_embed, const float* h,
                                 const float* cell, const float* coverage, const uint8_t* mask, float* probs,
                                 float* h_out, float* cell_out, float* att_cov, float* cov_out, long long* argmax,
                                 float* ctx, float* alpha, float* beta, float* gates, float* ctx12, int B, int Lt,
                                 int H, int E, int M, mmb_stream_t stream);

/* Backward of one decoder step (two launches: output layer + LSTM cell, then the attentions).
 * Inputs: the step's saved forward tensors (probs, h_out, cell_out, gates, alpha, beta, ctx12, and the
 * step inputs h, cell, coverage, proj_*, enc_*) and the incoming gradients d_probs (B,M), d_h_out,
 * d_cell_out (B,H), d_att_cov, d_cov_out (B,Lt).
 * Outputs
 *   d_h, d_cell (B,H), d_cov (B,Lt): gradients of the step inputs
 *   d_proj_a, d_proj_i (B,Lt,2H): ACCUMULATED in place (+=) across the steps of one sequence
 *   per-step rows for the deferred weight-gradient GEMMs:
 *     d_logits (B,M), d_gates (B,4H), d_ctx12 (B,2,2H) = d(loss)/d(c1), d(c2),
 *     d_pre (B,4,2H) = d tanh-argument sums [W2-side, W4-side, W_beta_1/2-side, W_beta_3/4-side]
 *   vec_acc (B,6,2H) += [dWc1, dWc2, dv1, dv2, dv_beta_1, dv_beta_2];  scal_acc (B,4) += their biases.
 */
MMB_API int mmb_decoder_step_bwd(const mmb_decoder_weights* w, const float* proj_a, const float* proj_i,
                                 const float* enc_a, const float* enc_i, const float* h, const float* cell,
                                 const float* coverage, const float* probs, const float* h_out, const float* cell_out,
                                 const float* gates, const float* alpha, const float* beta, const float* ctx12,
                                 const float* d_probs, const float* d_h_out, const float* d_cell_out,
                                 const float* d_att_cov, const float* d_cov_out, float* d_h, float* d_cell, float* d_cov,
                                 float* d_proj_a, float* d_proj_i, float* d_logits, float* d_gates, float* d_ctx12,
                                 float* d_pre, float* vec_acc, float* scal_acc, float* d_ctx, int B, int Lt, int H,
                                 int E, int M, mmb_stream_t stream);

/* --------------------------------------------------------------------------------------
 * masked_softmax over the last axis (attention.py:78-98): y = softmax(mask ? x : -1e30), or
 * log_softmax when log_mode != 0.  x, y (rows,n); mask (rows,n) bytes.  Backward returns
 * dx = mask * dsoftmax (the reference's d(mask*x)/dx = mask).
 */
MMB_API int mmb_masked_softmax_fwd(const float* x, const uint8_t* mask, float* y, long long rows, int n, int log_mode,
                                   mmb_stream_t stream);
MMB_API int mmb_masked_softmax_bwd(const float* y, const float* dy, const uint8_t* mask, float* dx, long long rows,
                                   int n, int log_mode, mmb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMBIDAF_B200_H_ */
