"""CPU oracle for the MMBiDAF hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, parameter-dict driven restatement (PyTorch CPU, any float
dtype) of the algorithm that the reference implements as ``nn.Module``s.  It is the
*checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing
under ``mmbidaf_b200/`` imports it, and the product path raises when the CUDA
library is missing rather than coming here.

Pinning.  The reference repository has no tests, golden vectors or fixtures for this
path, so the oracle is pinned differentially: the script
``tests/golden/make_golden.py`` imports the reference's own ``layers/*`` and
``models.py`` from ``/root/reference`` (in the build container, where that tree is
mounted), runs them on seeded inputs and commits inputs + outputs under
``tests/golden/*.pt``.  ``tests/test_oracle_golden.py`` replays every fixture through
this file.  Float arithmetic lives in PyTorch (third party, unpinned by the reference:
it ships no requirements file); fixtures were generated with torch 2.11.0 CPU fp32.

Every function cites the reference lines it follows (paths relative to the reference
root).  Parameters are plain dicts keyed by the reference's ``state_dict`` names so a
reference checkpoint can be fed in unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

NEG_FILL = -1e30  # layers/attention.py:94 uses -1e30, not -inf (quirk Q1)


def sub(params: Params, prefix: str) -> Params:
    """Select ``prefix.*`` entries of a flat state dict and strip the prefix."""
    plen = len(prefix) + 1
    return {k[plen:]: v for k, v in params.items() if k.startswith(prefix + ".")}


# ----------------------------------------------------------------------------------------
# masks (models.py:86-92, :116-123)
# ----------------------------------------------------------------------------------------
def length_mask(max_len: int, lengths: Sequence[int]) -> Tensor:
    """bool (B, max_len): position < length.  models.py:86-92 (get_mask)."""
    pos = torch.arange(max_len).unsqueeze(0)
    return pos < torch.tensor(list(lengths), dtype=torch.long).unsqueeze(1)


def decoder_mask(text_mask: Tensor, max_transcript_length: int) -> Tensor:
    """Pad the text mask with False up to M columns.  models.py:121-123."""
    b, lt = text_mask.shape
    pad = torch.zeros(b, max_transcript_length - lt, dtype=text_mask.dtype)
    return torch.cat([text_mask, pad], dim=1)


# ----------------------------------------------------------------------------------------
# masked softmax (layers/attention.py:78-98)
# ----------------------------------------------------------------------------------------
def masked_softmax(logits: Tensor, mask: Tensor, dim: int = -1, log_softmax: bool = False) -> Tensor:
    """softmax(mask*x + (1-mask)*-1e30).  The mask is cast to fp32 exactly as the
    reference does (attention.py:93), so an fp64 caller gets fp64 output via promotion."""
    m = mask.to(torch.float32)
    filled = m * logits + (1 - m) * NEG_FILL
    return torch.log_softmax(filled, dim) if log_softmax else torch.softmax(filled, dim)


# ----------------------------------------------------------------------------------------
# BiDAF attention (layers/attention.py:37-75)
# ----------------------------------------------------------------------------------------
def bidaf_similarity(p: Params, text: Tensor, modality: Tensor,
                     keep_text: Optional[Tensor] = None, keep_modality: Optional[Tensor] = None,
                     drop_prob: float = 0.0) -> Tensor:
    """Trilinear similarity S (B, Lc, Lq).  attention.py:56-75.

    ``keep_*`` are optional {0,1} dropout keep-masks; when given the inputs are scaled
    by keep/(1-p) as F.dropout does (attention.py:66-67, text first then modality)."""
    if keep_text is not None:
        text = text * keep_text / (1.0 - drop_prob)
    if keep_modality is not None:
        modality = modality * keep_modality / (1.0 - drop_prob)
    row_term = text @ p["text_weight"]                                  # (B, Lc, 1)    :70
    col_term = (modality @ p["modality_weight"]).transpose(1, 2)        # (B, 1, Lq)    :71
    cross = torch.einsum("bik,bjk->bij", text * p["text_modality_weight"], modality)   # :72
    return row_term + col_term + cross + p["bias"]                      # :73


def bidaf_attention(p: Params, text: Tensor, modality: Tensor, text_mask: Tensor, modality_mask: Tensor,
                    keep_text: Optional[Tensor] = None, keep_modality: Optional[Tensor] = None,
                    drop_prob: float = 0.0, reassociate: bool = False) -> Tensor:
    """(B, Lc, 4d) = [c, a, c*a, c*b].  attention.py:37-54.

    ``reassociate=False`` keeps the reference's (s1 s2^T) c order (attention.py:50);
    ``True`` uses s1 (s2^T c), the order the fused kernels use (differs ~1e-7 abs)."""
    s = bidaf_similarity(p, text, modality, keep_text, keep_modality, drop_prob)
    s1 = masked_softmax(s, modality_mask.unsqueeze(1), dim=2)           # :43
    s2 = masked_softmax(s, text_mask.unsqueeze(2), dim=1)               # :44
    a = torch.bmm(s1, modality)                                         # :47
    if reassociate:
        b = torch.bmm(s1, torch.bmm(s2.transpose(1, 2), text))
    else:
        b = torch.bmm(torch.bmm(s1, s2.transpose(1, 2)), text)          # :50
    return torch.cat([text, a, text * a, text * b], dim=2)              # :52


# ----------------------------------------------------------------------------------------
# Embedding + highway (layers/encoding.py:19-59)
# ----------------------------------------------------------------------------------------
def highway(p: Params, x: Tensor, num_layers: int = 2) -> Tensor:
    """encoding.py:52-59."""
    for k in range(num_layers):
        g = torch.sigmoid(x @ p[f"gates.{k}.weight"].t() + p[f"gates.{k}.bias"])
        t = torch.relu(x @ p[f"transforms.{k}.weight"].t() + p[f"transforms.{k}.bias"])
        x = g * t + (1 - g) * x
    return x


def embedding(p: Params, x: Tensor, keep: Optional[Tensor] = None, drop_prob: float = 0.0) -> Tensor:
    """dropout -> Linear(E->H, no bias) -> 2 highway layers.  encoding.py:25-30."""
    if keep is not None:
        x = x * keep / (1.0 - drop_prob)
    return highway(sub(p, "hwy"), x @ p["proj.weight"].t())


# ----------------------------------------------------------------------------------------
# Length-aware bidirectional LSTM (layers/encoding.py:76-108; torch.nn.LSTM semantics)
# ----------------------------------------------------------------------------------------
def lstm_cell(x_gates: Tensor, h: Tensor, c: Tensor, w_hh: Tensor) -> Tuple[Tensor, Tensor]:
    """One LSTM step given the input-side pre-activations (bias already added).
    PyTorch gate order i, f, g, o."""
    z = x_gates + h @ w_hh.t()
    i, f, g, o = z.chunk(4, dim=-1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c = f * c + i * g
    return o * torch.tanh(c), c


def lstm_direction(x: Tensor, lengths: Sequence[int], w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                   reverse: bool) -> Tuple[Tensor, Tensor]:
    """One direction of one layer over padded (B, L, in) with per-sample lengths.

    Returns (out (B, L, H) with exact zeros past each length -- pad_packed_sequence
    semantics, encoding.py:99 -- and the final hidden state (B, H) taken at each
    sample's own last valid step)."""
    bsz, max_len, _ = x.shape
    hid = w_hh.shape[1]
    xg = x @ w_ih.t() + b_ih + b_hh
    out = x.new_zeros(bsz, max_len, hid)
    h_last = x.new_zeros(bsz, hid)
    for b in range(bsz):
        n = int(lengths[b])
        h = x.new_zeros(hid)
        c = x.new_zeros(hid)
        steps = range(n - 1, -1, -1) if reverse else range(n)
        for t in steps:
            h, c = lstm_cell(xg[b, t], h, c, w_hh)
            out[b, t] = h
        h_last[b] = h
    return out, h_last


def sort_order(lengths: Sequence[int]) -> Tensor:
    """Descending-length permutation exactly as encoding.py:85,91 computes it (a CPU
    float tensor sorted with torch.sort, whose tie order we inherit by calling it)."""
    return torch.Tensor(list(lengths)).sort(0, descending=True)[1]


def rnn_encoder(p: Params, x: Tensor, lengths: Sequence[int], num_layers: int,
                keep_between: Optional[List[Tensor]] = None, keep_out: Optional[Tensor] = None,
                drop_prob: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Explicit-loop restatement of RNNEncoder.forward (encoding.py:83-108).

    ``p`` holds ``rnn.weight_ih_l{k}[_reverse]`` etc.  Returns (out (B, L, 2H) in
    batch order, h_n (B, 2*layers, H) in DESCENDING-LENGTH order: the reference
    un-sorts ``x`` but not ``x_hidden`` (quirk Q3, encoding.py:99-106)).
    ``keep_between[k]`` is the keep-mask applied to layer k's output before layer k+1
    (nn.LSTM inter-layer dropout); ``keep_out`` the one of encoding.py:104.  Both are
    expressed in batch order."""
    layer_in = x
    finals = []
    for k in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            o, h_last = lstm_direction(layer_in, lengths,
                                       p[f"rnn.weight_ih_l{k}{suffix}"], p[f"rnn.weight_hh_l{k}{suffix}"],
                                       p[f"rnn.bias_ih_l{k}{suffix}"], p[f"rnn.bias_hh_l{k}{suffix}"], rev)
            outs.append(o)
            finals.append(h_last)
        layer_in = torch.cat(outs, dim=2)
        if k + 1 < num_layers and keep_between is not None:
            layer_in = layer_in * keep_between[k] / (1.0 - drop_prob)
    out = layer_in
    if keep_out is not None:
        out = out * keep_out / (1.0 - drop_prob)
    h_n = torch.stack(finals, dim=1)                      # (B, 2*layers, H), batch order
    return out, h_n[sort_order(lengths)]                  # Q3: left in sorted order


def rnn_encoder_aten(p: Params, x: Tensor, lengths: Sequence[int], num_layers: int) -> Tuple[Tensor, Tensor]:
    """Same result as :func:`rnn_encoder` (no dropout) through torch's own packed-sequence
    LSTM, i.e. the library routine the reference calls (encoding.py:91-101).  Used where
    the oracle is *timed* as the CPU baseline, so that the port does not run slower than
    the reference merely because of Python loops."""
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    flat = []
    for k in range(num_layers):
        for suffix in ("", "_reverse"):
            flat += [p[f"rnn.weight_ih_l{k}{suffix}"], p[f"rnn.weight_hh_l{k}{suffix}"],
                     p[f"rnn.bias_ih_l{k}{suffix}"], p[f"rnn.bias_hh_l{k}{suffix}"]]
    hid = flat[1].shape[1]
    order = sort_order(lengths)
    sorted_len = torch.Tensor(list(lengths))[order]
    packed = pack_padded_sequence(x[order], sorted_len, batch_first=True)
    zeros = x.new_zeros(2 * num_layers, x.shape[0], hid)
    data, h_n, _ = torch._VF.lstm(packed.data, packed.batch_sizes, (zeros, zeros), flat, True,
                                  num_layers, 0.0, False, True)
    out, _ = pad_packed_sequence(
        torch.nn.utils.rnn.PackedSequence(data, packed.batch_sizes, None, None),
        batch_first=True, total_length=x.shape[1])
    inverse = order.sort(0)[1]
    return out[inverse], h_n.transpose(0, 1)


# ----------------------------------------------------------------------------------------
# Multimodal attention decoder step (layers/attention.py:145-186)
# ----------------------------------------------------------------------------------------
def _linear(p: Params, name: str, x: Tensor) -> Tensor:
    return x @ p[f"{name}.weight"].t() + p[f"{name}.bias"]


def decoder_step(p: Params, sent_embed: Tensor, h: Tensor, cell: Tensor, enc_audio: Tensor, enc_image: Tensor,
                 coverage: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """One step.  Shapes as the reference: sent_embed (B,1,E), h (B,1,H), cell (1,B,H),
    enc_* (B,Lt,2H), coverage (B,Lt,1), mask (B,M) -> (probs (B,M), h' (B,1,H),
    cell' (1,B,H), att_cov (B,Lt,1), coverage' (B,Lt,1))."""
    def additive(enc, w_enc, w_h, w_cov, v):                          # :147-150 / :153-156
        e = _linear(p, v, torch.tanh(_linear(p, w_enc, enc) + _linear(p, w_h, h) + _linear(p, w_cov, coverage)))
        alpha = torch.softmax(e, dim=1)                               # un-masked on purpose (quirk Q2)
        return alpha, (alpha * enc).sum(dim=1)

    alpha1, ctx1 = additive(enc_audio, "W1", "W2", "Wc1", "v1")
    alpha2, ctx2 = additive(enc_image, "W3", "W4", "Wc2", "v2")
    eb1 = _linear(p, "v_beta_1", torch.tanh(_linear(p, "W_beta_1", ctx1.unsqueeze(1)) + _linear(p, "W_beta_2", h)))   # :161
    eb2 = _linear(p, "v_beta_2", torch.tanh(_linear(p, "W_beta_3", ctx2.unsqueeze(1)) + _linear(p, "W_beta_4", h)))   # :162
    beta = torch.softmax(torch.cat([eb1, eb2], dim=1), dim=1)         # (B,2,1)      :163-164
    ctx = (torch.stack([ctx1, ctx2], dim=1) * beta).sum(dim=1)        # (B,2H)       :165-166
    att_cov = torch.bmm(torch.cat([alpha1, alpha2], dim=2), beta)     # (B,Lt,1)     :167
    coverage = coverage + att_cov                                     #              :177
    x = torch.cat([ctx, sent_embed.squeeze(1)], dim=1)                # (B, 2H+E)    :179
    xg = x @ p["lstm.weight_ih_l0"].t() + p["lstm.bias_ih_l0"] + p["lstm.bias_hh_l0"]
    h_new, c_new = lstm_cell(xg, h.squeeze(1), cell.squeeze(0), p["lstm.weight_hh_l0"])    # :181
    probs = masked_softmax(_linear(p, "out", h_new), mask)            #              :184
    return probs, h_new.unsqueeze(1), c_new.unsqueeze(0), att_cov, coverage


# ----------------------------------------------------------------------------------------
# Whole model (models.py:94-206).  The frozen ResNet (encoding.py:111-154) is out of
# scope: images arrive as (B, Li, E_img) feature rows (north_star: "image 1000-d").
# ----------------------------------------------------------------------------------------
def mmbidaf_forward(p: Params, text: Tensor, text_len: Sequence[int], audio: Tensor, audio_len: Sequence[int],
                    image_feat: Tensor, image_len: Sequence[int], targets: Tensor, max_dec_len: int,
                    max_transcript_length: int, training: bool, fast_lstm: bool = False
                    ) -> Tuple[Tensor, Tensor]:
    """Returns (out_distributions (B, T, M), loss).  Dropout-free (drop_prob 0 or eval)."""
    enc = rnn_encoder_aten if fast_lstm else rnn_encoder
    t_emb = embedding(sub(p, "emb"), text)                                            # :95
    t_enc, _ = enc(sub(p, "text_enc"), t_emb, text_len, 1)                            # :97
    a_enc, _ = enc(sub(p, "audio_enc"), embedding(sub(p, "a_emb"), audio), audio_len, 1)      # :100-102
    i_enc, _ = enc(sub(p, "image_enc"), embedding(sub(p, "i_emb"), image_feat), image_len, 1)  # :111-113
    t_mask = length_mask(text.shape[1], text_len)                                     # :116
    a_mask = length_mask(audio.shape[1], audio_len)
    i_mask = length_mask(image_feat.shape[1], image_len)
    d_mask = decoder_mask(t_mask, max_transcript_length)                              # :121-123
    att_a = bidaf_attention(sub(p, "bidaf_att_audio"), t_enc, a_enc, t_mask, a_mask)  # :131
    att_i = bidaf_attention(sub(p, "bidaf_att_image"), t_enc, i_enc, t_mask, i_mask)  # :132
    mod_a, hid_a = enc(sub(p, "mod_t_a"), att_a, text_len, 2)                         # :134
    mod_i, hid_i = enc(sub(p, "mod_t_i"), att_i, text_len, 2)                         # :135
    h = (hid_a.sum(1) + hid_i.sum(1)).unsqueeze(1)                                    # :143 (Q3: sorted-order rows)
    bsz, lt = text.shape[0], text.shape[1]
    cell = text.new_zeros(1, bsz, h.shape[-1])                                        # :145
    dec_in = text.new_zeros(bsz, 1, text.shape[-1])                                   # :147
    cov = text.new_zeros(bsz, lt, 1)                                                  # :149
    dp = sub(p, "multimodal_att_decoder")
    rows = torch.arange(bsz)
    loss = text.new_zeros(())
    dists = []
    steps = targets.shape[1] if training else max_dec_len                             # :162 / :182
    att_cov = None
    for t in range(steps):
        probs, h, cell, att_cov, cov = decoder_step(dp, dec_in, h, cell, mod_a, mod_i, cov, d_mask)
        tgt = targets[:, t].reshape(bsz).long()                                       # int(tensor) :168/:188
        loss = loss - torch.log(probs[rows, tgt] + 1e-12).sum()                       # :170 (summed over batch, Q6)
        nxt = tgt if training else probs.max(dim=1)[1]                                # :173 / :184,:193
        dec_in = text[rows, nxt].unsqueeze(1)
        dists.append(probs)
        if training:
            loss = loss + torch.min(att_cov, cov).sum()                               # :177-178 (every step, Q5)
    if not training:
        loss = loss + torch.min(att_cov, cov).sum()                                   # :197-198 (once)
    return torch.stack(dists).transpose(0, 1), loss / steps                           # :179/:199, :205


def greedy_indices(out_distributions: Tensor, text_len: int) -> List[int]:
    """Selected sentence indices of one video.  evaluate.py:185-202 with the disk look-up of
    ``get_source_sentence`` (:236-259) reduced to what it decides: the transcript has ``text_len - 1``
    sentences (the text length includes the EOS row, datasets.py:70); stop when the arg-max equals the EOS
    row ``text_len - 1`` (:188, :253-254), SKIP an index beyond it (:255-256 returns None, :195), keep the rest.
    Pinned by tests/golden/greedy_search.pt (made by the reference's own evaluate.py)."""
    picked = []
    for row in out_distributions.detach().cpu().numpy():
        k = int(row.argmax())
        if k == text_len - 1:
            break
        if k > text_len - 1:
            continue
        picked.append(k)
    return picked


# ----------------------------------------------------------------------------------------
# Parameter construction with the reference's names and shapes (models.py:29-83)
# ----------------------------------------------------------------------------------------
def param_shapes(hidden: int, e_text: int, e_audio: int, e_image: int, max_transcript_length: int
                 ) -> Dict[str, Tuple[int, ...]]:
    h, d = hidden, 2 * hidden
    shapes: Dict[str, Tuple[int, ...]] = {}
    for name, e in (("emb", e_text), ("a_emb", e_audio), ("i_emb", e_image)):
        shapes[f"{name}.proj.weight"] = (h, e)
        for k in range(2):
            for part in ("transforms", "gates"):
                shapes[f"{name}.hwy.{part}.{k}.weight"] = (h, h)
                shapes[f"{name}.hwy.{part}.{k}.bias"] = (h,)

    def lstm(prefix, in0, layers, bidir=True):
        for k in range(layers):
            fan_in = in0 if k == 0 else (2 * h if bidir else h)
            for suffix in (("", "_reverse") if bidir else ("",)):
                shapes[f"{prefix}.weight_ih_l{k}{suffix}"] = (4 * h, fan_in)
                shapes[f"{prefix}.weight_hh_l{k}{suffix}"] = (4 * h, h)
                shapes[f"{prefix}.bias_ih_l{k}{suffix}"] = (4 * h,)
                shapes[f"{prefix}.bias_hh_l{k}{suffix}"] = (4 * h,)

    for name in ("text_enc", "audio_enc", "image_enc"):
        lstm(f"{name}.rnn", h, 1)
    for name in ("bidaf_att_audio", "bidaf_att_image"):
        shapes[f"{name}.text_weight"] = (d, 1)
        shapes[f"{name}.modality_weight"] = (d, 1)
        shapes[f"{name}.text_modality_weight"] = (1, 1, d)
        shapes[f"{name}.bias"] = (1,)
    for name in ("mod_t_a", "mod_t_i"):
        lstm(f"{name}.rnn", 8 * h, 2)
    dec = "multimodal_att_decoder"
    for name, (o, i) in {"W1": (d, d), "W2": (d, h), "Wc1": (d, 1), "v1": (1, d),
                         "W3": (d, d), "W4": (d, h), "Wc2": (d, 1), "v2": (1, d),
                         "W_beta_1": (d, d), "W_beta_2": (d, h), "W_beta_3": (d, d), "W_beta_4": (d, h),
                         "v_beta_1": (1, d), "v_beta_2": (1, d), "out": (max_transcript_length, h)}.items():
        shapes[f"{dec}.{name}.weight"] = (o, i)
        shapes[f"{dec}.{name}.bias"] = (o,)
    lstm(f"{dec}.lstm", e_text + d, 1, bidir=False)
    return shapes


def make_params(hidden: int, e_text: int, e_audio: int, e_image: int, max_transcript_length: int,
                seed: int = 224, dtype=torch.float32) -> Params:
    """Deterministic parameters (uniform +-1/sqrt(fan_in); sorted-name order) so fixtures
    need only store the seed.  Not the reference's init scheme -- values are loaded INTO
    the reference by ``load_state_dict`` when fixtures are generated."""
    gen = torch.Generator().manual_seed(seed)
    out: Params = {}
    shapes = param_shapes(hidden, e_text, e_audio, e_image, max_transcript_length)
    for name in sorted(shapes):
        shape = shapes[name]
        fan_in = shape[-1] if len(shape) > 1 else hidden
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        out[name] = ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return out
