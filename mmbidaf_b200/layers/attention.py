"""Attention blocks of the MMBiDAF hot path on B200.

Same public surface as the reference's layers/attention.py: ``BiDAFAttention``, ``masked_softmax``,
``MultimodalAttentionDecoder`` with identical constructor / ``forward`` signatures, parameter names and
shapes.  The arithmetic runs in the fused sm_100a kernels of csrc/ (bidaf_*.cu, decoder_fwd.cu,
masked_softmax.cu); nothing here falls back to a CPU or library implementation of those ops.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as Fn
from .. import ops

__all__ = ["BiDAFAttention", "masked_softmax", "MultimodalAttentionDecoder"]


def _keep_mask(x, drop_prob):
    """Bernoulli(1-p) keep mask as uint8, drawn with torch's generator on x's device."""
    return torch.empty(x.shape, dtype=torch.uint8, device=x.device).bernoulli_(1.0 - drop_prob)


class BiDAFAttention(nn.Module):
    """Bidirectional text<->modality attention (reference attention.py:9-75).

    forward(text (B,Lc,d), modality (B,Lq,d), text_mask (B,Lc), modality_mask (B,Lq)) -> (B,Lc,4d)
    = [text, a, text*a, text*b].  ``precision`` selects the contraction tier of the fused kernel
    ("fp32": rel <= 1e-5; "bf16": tcgen05 tensor cores, rel <= 2e-2)."""

    precision = "fp32"

    def __init__(self, hidden_size, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob
        self.text_weight = nn.Parameter(torch.zeros(hidden_size, 1))
        self.modality_weight = nn.Parameter(torch.zeros(hidden_size, 1))
        self.text_modality_weight = nn.Parameter(torch.zeros(1, 1, hidden_size))
        for weight in (self.text_weight, self.modality_weight, self.text_modality_weight):
            nn.init.xavier_uniform_(weight)
        self.bias = nn.Parameter(torch.zeros(1))

    def _dropout_masks(self, text, modality):
        if not self.training or self.drop_prob <= 0:
            return None, None, 1.0
        # same draw order as the reference: text first, then modality (attention.py:66-67)
        return _keep_mask(text, self.drop_prob), _keep_mask(modality, self.drop_prob), 1.0 / (1.0 - self.drop_prob)

    def forward(self, text, modality, text_mask, modality_mask):
        if not text.is_cuda:
            raise RuntimeError("mmbidaf_b200.layers.BiDAFAttention runs on a B200 only (no CPU fallback)")
        keep_c, keep_q, scale = self._dropout_masks(text, modality)
        prec = ops.PREC_BF16 if self.precision == "bf16" else ops.PREC_FP32
        return Fn.bidaf_attention(text, modality, text_mask, modality_mask, self.text_weight, self.modality_weight,
                                  self.text_modality_weight, self.bias, keep_c, keep_q, scale, prec)

    def get_similarity_matrix(self, text, modality):
        """Materialised S (B,Lc,Lq).  Kept for API compatibility (attention.py:56-75); ``forward`` never
        calls it -- the fused kernel builds S tile by tile on chip."""
        text = F.dropout(text, self.drop_prob, self.training)
        modality = F.dropout(modality, self.drop_prob, self.training)
        s = torch.baddbmm(text @ self.text_weight + (modality @ self.modality_weight).transpose(1, 2),
                          text * self.text_modality_weight, modality.transpose(1, 2))
        return s + self.bias


def masked_softmax(logits, mask, dim=-1, log_softmax=False):
    """softmax(mask*x + (1-mask)*-1e30) along ``dim`` (reference attention.py:78-98), as a CUDA kernel."""
    return Fn.masked_softmax(logits, mask, dim, log_softmax)


class MultimodalAttentionDecoder(nn.Module):
    """Pointer-style decoder step with two coverage attentions (reference attention.py:100-186).

    forward(sent_embed (B,1,E), decoder_hidden (B,1,H), decoder_cell_state (1,B,H), text_audio_enc_out
    (B,Lt,2H), text_img_enc_out (B,Lt,2H), coverage_vec (B,Lt,1), mask (B,M)) -> (final_out (B,M),
    decoder_hidden (B,1,H), decoder_cell_state (1,B,H), att_cov_dist (B,Lt,1), coverage_vec (B,Lt,1)).

    ``W1(text_audio_enc_out)`` and ``W3(text_img_enc_out)`` do not change between steps; they are computed
    once per pair of encoder tensors and cached on the tensors' identity."""

    def __init__(self, text_embedding_size, hidden_size, output_size, num_layers=1, dropout=0.1):
        super().__init__()
        self.text_embedding_size = text_embedding_size
        self.hidden_size = hidden_size
        self.output_size = output_size
        self.num_layers = num_layers
        self.dropout = dropout
        h2 = 2 * hidden_size
        self.W1 = nn.Linear(h2, h2)
        self.W2 = nn.Linear(hidden_size, h2)
        self.Wc1 = nn.Linear(1, h2)
        self.v1 = nn.Linear(h2, 1)
        self.tanh = nn.Tanh()
        self.W3 = nn.Linear(h2, h2)
        self.W4 = nn.Linear(hidden_size, h2)
        self.Wc2 = nn.Linear(1, h2)
        self.v2 = nn.Linear(h2, 1)
        self.W_beta_1 = nn.Linear(h2, h2)
        self.W_beta_2 = nn.Linear(hidden_size, h2)
        self.W_beta_3 = nn.Linear(h2, h2)
        self.W_beta_4 = nn.Linear(hidden_size, h2)
        self.v_beta_1 = nn.Linear(h2, 1)
        self.v_beta_2 = nn.Linear(h2, 1)
        self.lstm = nn.LSTM(text_embedding_size + h2, hidden_size, num_layers, batch_first=True)
        self.out = nn.Linear(hidden_size, output_size)
        self.softmax = nn.Softmax()
        self._cache = None

    # ---- step-invariant projections --------------------------------------------------------------------
    def _projections(self, enc_a, enc_i):
        c = self._cache
        if c is not None and c[0]() is enc_a and c[1]() is enc_i and c[2] == (enc_a._version, enc_i._version,
                                                                             torch.is_grad_enabled()):
            return c[3], c[4]
        proj_a, proj_i = self.W1(enc_a), self.W3(enc_i)
        self._cache = (weakref.ref(enc_a), weakref.ref(enc_i), (enc_a._version, enc_i._version, torch.is_grad_enabled()),
                       proj_a, proj_i)
        return proj_a, proj_i

    def _weight_struct(self):
        g = lambda lin: (lin.weight.detach().contiguous(), lin.bias.detach().contiguous())
        t = {}
        for name, lin in (("W2", self.W2), ("Wc1", self.Wc1), ("v1", self.v1), ("W4", self.W4), ("Wc2", self.Wc2),
                          ("v2", self.v2)):
            t[name], t[("b" + name[1:]) if name[0] == "W" else name + "b"] = g(lin)
        for k, lin in (("1", self.W_beta_1), ("2", self.W_beta_2), ("3", self.W_beta_3), ("4", self.W_beta_4)):
            t["Wb" + k], t["bb" + k] = g(lin)
        t["vb1"], t["vb1b"] = g(self.v_beta_1)
        t["vb2"], t["vb2b"] = g(self.v_beta_2)
        t["lstm_w_ih"], t["lstm_w_hh"] = self.lstm.weight_ih_l0.detach().contiguous(), self.lstm.weight_hh_l0.detach().contiguous()
        t["lstm_b_ih"], t["lstm_b_hh"] = self.lstm.bias_ih_l0.detach().contiguous(), self.lstm.bias_hh_l0.detach().contiguous()
        t["out_w"], t["out_b"] = g(self.out)
        return ops.decoder_weights(t), t

    def forward(self, sent_embed, decoder_hidden, decoder_cell_state, text_audio_enc_out, text_img_enc_out,
                coverage_vec, mask):
        if not text_audio_enc_out.is_cuda:
            raise RuntimeError("mmbidaf_b200.layers.MultimodalAttentionDecoder runs on a B200 only (no CPU fallback)")
        if self.num_layers != 1:
            raise RuntimeError("MultimodalAttentionDecoder: only num_layers=1 is supported (the reference model uses 1)")
        proj_a, proj_i = self._projections(text_audio_enc_out, text_img_enc_out)
        needs_grad = torch.is_grad_enabled() and (
            proj_a.requires_grad or decoder_hidden.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            return self._step_autograd(sent_embed, decoder_hidden, decoder_cell_state, text_audio_enc_out,
                                       text_img_enc_out, proj_a, proj_i, coverage_vec, mask)
        B, Lt, _ = text_audio_enc_out.shape
        w, keep_alive = self._weight_struct()
        probs, h, cell, att, cov, _, _ = ops.decoder_step_fwd(
            w, proj_a.contiguous(), proj_i.contiguous(), text_audio_enc_out.contiguous(), text_img_enc_out.contiguous(),
            sent_embed.reshape(B, -1).contiguous(), decoder_hidden.reshape(B, -1).contiguous(),
            decoder_cell_state.reshape(B, -1).contiguous(), coverage_vec.reshape(B, Lt).contiguous(),
            ops._u8(mask), self.output_size)
        del keep_alive
        return probs, h.unsqueeze(1), cell.unsqueeze(0), att.unsqueeze(2), cov.unsqueeze(2)

    def _step_autograd(self, sent_embed, h, cell, enc_a, enc_i, proj_a, proj_i, coverage, mask):
        """Training step (INTERIM, round 1): same arithmetic as the fused kernels, expressed with cuBLAS /
        ATen ops on the GPU so that autograd supplies the backward pass; the step-invariant projections are
        still hoisted.  A fused backward kernel replaces this next ("decoder backward")."""
        e1 = self.v1(torch.tanh(proj_a + self.W2(h) + self.Wc1(coverage)))
        a1 = F.softmax(e1, dim=1)
        c1 = (a1 * enc_a).sum(dim=1)
        e2 = self.v2(torch.tanh(proj_i + self.W4(h) + self.Wc2(coverage)))
        a2 = F.softmax(e2, dim=1)
        c2 = (a2 * enc_i).sum(dim=1)
        eb1 = self.v_beta_1(torch.tanh(self.W_beta_1(c1.unsqueeze(1)) + self.W_beta_2(h)))
        eb2 = self.v_beta_2(torch.tanh(self.W_beta_3(c2.unsqueeze(1)) + self.W_beta_4(h)))
        beta = F.softmax(torch.cat((eb1, eb2), dim=1), dim=1)
        c3 = (torch.stack((c1, c2), dim=1) * beta).sum(dim=1)
        att = torch.bmm(torch.cat((a1, a2), dim=2), beta)
        coverage = coverage + att
        x = torch.cat((c3, sent_embed.squeeze(1)), dim=1)
        gates = F.linear(x, self.lstm.weight_ih_l0, self.lstm.bias_ih_l0) + \
            F.linear(h.squeeze(1), self.lstm.weight_hh_l0, self.lstm.bias_hh_l0)
        gi, gf, gg, go = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(gf) * cell.squeeze(0) + torch.sigmoid(gi) * torch.tanh(gg)
        h_new = torch.sigmoid(go) * torch.tanh(c_new)
        logits = self.out(h_new)
        probs = F.softmax(torch.where(mask.bool(), logits, logits.new_full((), -1e30)), dim=-1)
        return probs, h_new.unsqueeze(1), c_new.unsqueeze(0), att, coverage
