"""autograd glue: torch.autograd.Function wrappers whose forward AND backward run the library's kernels.

GEMMs that are plain library products (the LSTM input projection x W_ih^T and its transposes in the
backward pass) go through torch.mm / cuBLAS with TF32 off; everything recurrent, soft-maxed or fused is
a hand-written kernel reached through ops.py.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


class _LstmLayer(torch.autograd.Function):
    """One (bi)directional LSTM layer over padded (B, L, in) with per-sample lengths."""

    @staticmethod
    def forward(ctx, x, lengths, order, *weights):
        # weights = (w_ih, w_hh, b_ih, b_hh) per direction
        ndir = len(weights) // 4
        B, L, fan_in = x.shape
        H = weights[1].shape[1]
        w_ih = torch.cat([weights[4 * d] for d in range(ndir)], dim=0)                       # (ndir*4H, in)
        w_hh = torch.stack([weights[4 * d + 1] for d in range(ndir)], dim=0).contiguous()    # (ndir, 4H, H)
        bias = torch.cat([weights[4 * d + 2] + weights[4 * d + 3] for d in range(ndir)], dim=0)
        x2d = x.reshape(B * L, fan_in)
        gates = torch.addmm(bias, x2d, w_ih.t())                                            # plain GEMM (cuBLAS)
        save = any(ctx.needs_input_grad)
        out, h_n, c_n, cell = ops.lstm_layer_fwd(gates, w_hh, lengths, order, B, L, H, ndir, save)
        if save:
            ctx.save_for_backward(x2d, gates, cell, out, w_ih, w_hh, lengths, order)
            ctx.dims = (B, L, H, ndir, fan_in)
        ctx.mark_non_differentiable(c_n)
        return out, h_n, c_n

    @staticmethod
    def backward(ctx, d_out, d_h_n, _d_c_n):
        x2d, gates, cell, out, w_ih, w_hh, lengths, order = ctx.saved_tensors
        B, L, H, ndir, fan_in = ctx.dims
        if d_out is None:
            d_out = torch.zeros_like(out)
        da = ops.lstm_layer_bwd(gates, cell, w_hh, lengths, order, d_out, d_h_n, None, B, L, H, ndir)
        da2d = da.view(B * L, ndir * 4 * H)
        dx = (da2d @ w_ih).view(B, L, fan_in) if ctx.needs_input_grad[0] else None
        dw_ih = da2d.t() @ x2d                                                               # (ndir*4H, in)
        db = da2d.sum(dim=0)
        grads = []
        zero = out.new_zeros(B, 1, H)
        for d in range(ndir):
            h_dir = out[:, :, d * H:(d + 1) * H]
            # state that fed step t: h[t-1] going forward, h[t+1] going backward (zero at the start)
            h_prev = torch.cat([zero, h_dir[:, :-1]], dim=1) if d == 0 else torch.cat([h_dir[:, 1:], zero], dim=1)
            da_dir = da2d[:, d * 4 * H:(d + 1) * 4 * H]
            dw_hh = da_dir.t() @ h_prev.reshape(B * L, H)
            b_d = db[d * 4 * H:(d + 1) * 4 * H]
            grads += [dw_ih[d * 4 * H:(d + 1) * 4 * H], dw_hh, b_d, b_d]
        return (dx, None, None, *grads)


def lstm_layer(x: torch.Tensor, lengths: torch.Tensor, order: Optional[torch.Tensor], weights
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x (B,L,in) fp32 CUDA; lengths/order int32 CUDA; weights = [w_ih, w_hh, b_ih, b_hh] * ndir.
    Returns (out (B,L,ndir*H), h_n (B,ndir,H)) in batch order."""
    out, h_n, _ = _LstmLayer.apply(x.contiguous(), lengths, order, *weights)
    return out, h_n
