"""Thin Python wrappers over the C ABI: allocate outputs with torch, pass raw pointers + stream.

Every function here launches hand-written sm_100a kernels from libmmbidaf_b200.so and raises
if the library or a B200 is missing.  ``launch_count`` counts kernel launches issued through
this module (bench.py reports it as ``gpu_launches``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

PREC_FP32, PREC_BF16 = 0, 1
launch_count = 0


def _count(n: int) -> None:
    global launch_count
    launch_count += n


def _u8(mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return mask.to(torch.uint8).contiguous()


def bidaf_fwd(text: torch.Tensor, modality: torch.Tensor, text_mask: torch.Tensor, modality_mask: torch.Tensor,
              w_text: torch.Tensor, w_modality: torch.Tensor, w_cross: torch.Tensor, bias: torch.Tensor,
              keep_text: Optional[torch.Tensor] = None, keep_modality: Optional[torch.Tensor] = None,
              keep_scale: float = 1.0, precision: int = PREC_FP32
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fused BiDAF forward (attention.py:37-75).  Returns (out (B,Lc,4d), q2c (B,Lq,d),
    lse_row (B,Lc), lse_col (B,Lq)); the last three are what the backward pass needs."""
    L = _lib.lib()
    assert text.dtype == torch.float32 and modality.dtype == torch.float32
    B, Lc, d = text.shape
    Lq = modality.shape[1]
    text, modality = text.contiguous(), modality.contiguous()
    tm, mm = _u8(text_mask.reshape(B, Lc)), _u8(modality_mask.reshape(B, Lq))
    kt, km = _u8(keep_text), _u8(keep_modality)
    wt, wm, wc = (w.detach().reshape(-1).contiguous() for w in (w_text, w_modality, w_cross))
    out = torch.empty(B, Lc, 4 * d, device=text.device, dtype=torch.float32)
    q2c = torch.empty(B, Lq, d, device=text.device, dtype=torch.float32)
    lse_row = torch.empty(B, Lc, device=text.device, dtype=torch.float32)
    lse_col = torch.empty(B, Lq, device=text.device, dtype=torch.float32)
    p = _lib.ptr
    ws_bytes = L.mmb_bidaf_workspace_bytes(B, Lc, Lq, d, int(precision), int(km is not None))
    ws = torch.empty(ws_bytes, device=text.device, dtype=torch.uint8) if ws_bytes else None
    _lib.check(L.mmb_bidaf_fwd(p(text), p(modality), p(tm), p(mm), p(wt), p(wm), p(wc), p(bias.detach().contiguous()),
                               p(kt), p(km), float(keep_scale), p(out), p(q2c), p(lse_row), p(lse_col), p(ws),
                               B, Lc, Lq, d, int(precision), _lib.stream()), "mmb_bidaf_fwd")
    _count(2 if precision == PREC_FP32 else 3)
    return out, q2c, lse_row, lse_col


def lstm_layer_fwd(gates: torch.Tensor, w_hh: torch.Tensor, lengths: torch.Tensor, order: Optional[torch.Tensor],
                   B: int, L: int, H: int, ndir: int, save: bool
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Persistent LSTM recurrence of one layer (encoding.py:96).  ``gates`` (B,L,ndir,4H) holds the input
    projection on entry and, when ``save``, the activated gates on exit.  Returns (out (B,L,ndir*H),
    h_n (B,ndir,H), c_n (B,ndir,H), cell (B,L,ndir,H) or None)."""
    lib = _lib.lib()
    assert gates.dtype == torch.float32 and gates.numel() == B * L * ndir * 4 * H
    assert lengths.dtype == torch.int32 and (order is None or order.dtype == torch.int32)
    dev = gates.device
    out = torch.empty(B, L, ndir * H, device=dev, dtype=torch.float32)
    h_n = torch.empty(B, ndir, H, device=dev, dtype=torch.float32)
    c_n = torch.empty(B, ndir, H, device=dev, dtype=torch.float32)
    cell = torch.empty(B, L, ndir, H, device=dev, dtype=torch.float32) if save else None
    p = _lib.ptr
    _lib.check(lib.mmb_bilstm_fwd(p(gates), p(w_hh), p(lengths), p(order), p(out), p(h_n), p(c_n), p(cell),
                                  B, L, H, ndir, int(save), _lib.stream()), "mmb_bilstm_fwd")
    _count(1)
    return out, h_n, c_n, cell


def lstm_layer_bwd(gates: torch.Tensor, cell: torch.Tensor, w_hh: torch.Tensor, lengths: torch.Tensor,
                   order: Optional[torch.Tensor], dout: torch.Tensor, dh_n: Optional[torch.Tensor],
                   dc_n: Optional[torch.Tensor], B: int, L: int, H: int, ndir: int) -> torch.Tensor:
    """BPTT of :func:`lstm_layer_fwd`; overwrites ``gates`` with d(pre-activation) and returns it."""
    lib = _lib.lib()
    p = _lib.ptr
    _lib.check(lib.mmb_bilstm_bwd(p(gates), p(cell), p(w_hh), p(lengths), p(order), p(dout.contiguous()),
                                  p(None if dh_n is None else dh_n.contiguous()),
                                  p(None if dc_n is None else dc_n.contiguous()), B, L, H, ndir, _lib.stream()),
               "mmb_bilstm_bwd")
    _count(1)
    return gates


def decoder_weights(params: dict) -> "_lib.DecoderWeights":
    """Pack device pointers of the decoder parameters (keys = _lib.DECODER_WEIGHT_FIELDS)."""
    w = _lib.DecoderWeights()
    for name in _lib.DECODER_WEIGHT_FIELDS:
        t = params[name]
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32, name
        setattr(w, name, t.data_ptr())
    return w


def decoder_step_fwd(w, proj_a, proj_i, enc_a, enc_i, sent_embed, h, cell, coverage, mask_u8, M: int,
                     want_argmax: bool = False, save: bool = False):
    """One fused decoder step (attention.py:145-186).  All tensors 2-D/3-D contiguous fp32 CUDA:
    proj_*/enc_* (B,Lt,2H), sent_embed (B,E), h/cell (B,H), coverage (B,Lt), mask_u8 (B,M) uint8.
    Returns (probs, h', cell', att_cov, coverage', argmax|None, saved) with saved = (ctx, alpha, beta, gates)."""
    import ctypes
    lib = _lib.lib()
    B, Lt, D = enc_a.shape
    H, E = D // 2, sent_embed.shape[1]
    dev = enc_a.device
    f32 = dict(device=dev, dtype=torch.float32)
    probs = torch.empty(B, M, **f32)
    h_out, cell_out = torch.empty(B, H, **f32), torch.empty(B, H, **f32)
    att_cov, cov_out = torch.empty(B, Lt, **f32), torch.empty(B, Lt, **f32)
    ctx = torch.empty(B, D, **f32)
    argmax = torch.empty(B, device=dev, dtype=torch.int64) if want_argmax else None
    alpha = torch.empty(B, 2, Lt, **f32) if save else None
    beta = torch.empty(B, 2, **f32) if save else None
    gates = torch.empty(B, 4 * H, **f32) if save else None
    ctx12 = torch.empty(B, 2, D, **f32) if save else None
    p = _lib.ptr
    _lib.check(lib.mmb_decoder_step_fwd(ctypes.addressof(w), p(proj_a), p(proj_i), p(enc_a), p(enc_i), p(sent_embed),
                                        p(h), p(cell), p(coverage), p(mask_u8), p(probs), p(h_out), p(cell_out),
                                        p(att_cov), p(cov_out), p(argmax), p(ctx), p(alpha), p(beta), p(gates),
                                        p(ctx12), B, Lt, H, E, M, _lib.stream()), "mmb_decoder_step_fwd")
    _count(3)
    return probs, h_out, cell_out, att_cov, cov_out, argmax, (ctx, alpha, beta, gates, ctx12)


def decoder_step_bwd(w, proj_a, proj_i, enc_a, enc_i, h, cell, coverage, probs, h_out, cell_out, gates, alpha, beta,
                     ctx12, d_probs, d_h_out, d_cell_out, d_att_cov, d_cov_out, d_proj_a, d_proj_i, vec_acc, scal_acc,
                     E: int, M: int):
    """Backward of one decoder step.  ``d_proj_*`` (B,Lt,2H), ``vec_acc`` (B,6,2H) and ``scal_acc`` (B,4) are
    accumulated in place.  Returns (d_h, d_cell, d_cov, d_logits, d_gates, d_ctx12, d_pre)."""
    import ctypes
    lib = _lib.lib()
    B, Lt, D = enc_a.shape
    H = D // 2
    f32 = dict(device=enc_a.device, dtype=torch.float32)
    d_h, d_cell, d_cov = torch.empty(B, H, **f32), torch.empty(B, H, **f32), torch.empty(B, Lt, **f32)
    d_logits, d_gates = torch.empty(B, M, **f32), torch.empty(B, 4 * H, **f32)
    d_ctx12, d_pre, d_ctx = torch.empty(B, 2, D, **f32), torch.empty(B, 4, D, **f32), torch.empty(B, D, **f32)
    p = _lib.ptr
    c = lambda t: None if t is None else t.contiguous()
    _lib.check(lib.mmb_decoder_step_bwd(ctypes.addressof(w), p(proj_a), p(proj_i), p(enc_a), p(enc_i), p(h), p(cell),
                                        p(coverage), p(probs), p(h_out), p(cell_out), p(gates), p(alpha), p(beta),
                                        p(ctx12), p(c(d_probs)), p(c(d_h_out)), p(c(d_cell_out)), p(c(d_att_cov)),
                                        p(c(d_cov_out)), p(d_h), p(d_cell), p(d_cov), p(d_proj_a), p(d_proj_i),
                                        p(d_logits), p(d_gates), p(d_ctx12), p(d_pre), p(vec_acc), p(scal_acc),
                                        p(d_ctx), B, Lt, H, E, M, _lib.stream()), "mmb_decoder_step_bwd")
    _count(2)
    return d_h, d_cell, d_cov, d_logits, d_gates, d_ctx12, d_pre


def masked_softmax_fwd(x2d: torch.Tensor, mask2d_u8: torch.Tensor, log_mode: bool) -> torch.Tensor:
    lib = _lib.lib()
    rows, n = x2d.shape
    y = torch.empty_like(x2d)
    _lib.check(lib.mmb_masked_softmax_fwd(_lib.ptr(x2d), _lib.ptr(mask2d_u8), _lib.ptr(y), rows, n, int(log_mode),
                                          _lib.stream()), "mmb_masked_softmax_fwd")
    _count(1)
    return y


def masked_softmax_bwd(y2d: torch.Tensor, dy2d: torch.Tensor, mask2d_u8: torch.Tensor, log_mode: bool) -> torch.Tensor:
    lib = _lib.lib()
    rows, n = y2d.shape
    dx = torch.empty_like(y2d)
    _lib.check(lib.mmb_masked_softmax_bwd(_lib.ptr(y2d), _lib.ptr(dy2d.contiguous()), _lib.ptr(mask2d_u8), _lib.ptr(dx),
                                          rows, n, int(log_mode), _lib.stream()), "mmb_masked_softmax_bwd")
    _count(1)
    return dx
