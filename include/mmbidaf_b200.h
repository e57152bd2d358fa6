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
 *   soft-max -- saved for mmb_bidaf_bwd.  d % 4 == 0, d <= 256.
 */
MMB_API int mmb_bidaf_fwd(const float* text, const float* modality, const uint8_t* text_mask, const uint8_t* modality_mask,
                  const float* w_text, const float* w_modality, const float* w_cross, const float* bias,
                  const uint8_t* keep_text, const uint8_t* keep_modality, float keep_scale,
                  float* out, float* q2c, float* lse_row, float* lse_col,
                  int B, int Lc, int Lq, int d, int precision, mmb_stream_t stream);

/* --------------------------------------------------------------------------------------
 * Length-aware LSTM recurrence of one layer, 1 or 2 directions.  Replaces the nn.LSTM call of
 * encoding.py:96 together with the sort / pack / pad / unsort gathers of encoding.py:91-101.
 *   gates (B,L,ndir,4H): on entry x W_ih^T + b_ih + b_hh (gate order i,f,g,o; a plain GEMM done
 *     by the caller); with save != 0 it holds the ACTIVATED gates on exit (for the backward pass)
 *   w_hh (ndir,4H,H)   lengths (B) int32   order (B) int32 permutation, longest first, or NULL
 *   out (B,L,ndir*H): hidden states, exact zeros past each length (pad_packed_sequence)
 *   h_n, c_n (B,ndir,H): state after each sample's last valid step, in BATCH order
 *   cell (B,L,ndir,H): cell states, written when save != 0 (may be NULL otherwise).  H <= 128.
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

#ifdef __cplusplus
}
#endif
#endif /* MMBIDAF_B200_H_ */
