"""The library bar on the same B200: torch-eager CUDA of the reference's BiDAF attention (cuBLAS + ATen, TF32 off
and on) and cuDNN's nn.LSTM on a PackedSequence, timed with CUDA events next to our kernels.
    python tools/eager_baselines.py
"""
import os
import sys

import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import functional as F, ops  # noqa: E402

dev = "cuda"


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3          # us


def masked_softmax(logits, mask, dim):                # layers/attention.py:78-98
    mask = mask.type(torch.float32)
    return torch.softmax(mask * logits + (1 - mask) * -1e30, dim)


def eager_bidaf(c, q, cm, qm, w_c, w_q, w_cq, bias):   # layers/attention.py:37-75, eval mode
    B, Lc, _ = c.shape
    Lq = q.size(1)
    s0 = torch.matmul(c, w_c).expand([-1, -1, Lq])
    s1 = torch.matmul(q, w_q).transpose(1, 2).expand([-1, Lc, -1])
    s2 = torch.matmul(c * w_cq, q.transpose(1, 2))
    s = s0 + s1 + s2 + bias
    p = masked_softmax(s, qm.view(B, 1, Lq), 2)
    r = masked_softmax(s, cm.view(B, Lc, 1), 1)
    a = torch.bmm(p, q)
    b = torch.bmm(torch.bmm(p, r.transpose(1, 2)), c)
    return torch.cat([c, a, c * a, c * b], dim=2)


# ---- BiDAF at BASELINE config 2 ------------------------------------------------------------------------------------------
B, Lc, Lq, d = 64, 512, 256, 200
gen = torch.Generator().manual_seed(224)
c = torch.randn(B, Lc, d, generator=gen).to(dev)
q = torch.randn(B, Lq, d, generator=gen).to(dev)
cm = (torch.arange(Lc).unsqueeze(0) < torch.randint(Lc // 2, Lc + 1, (B, 1), generator=gen)).to(dev)
qm = (torch.arange(Lq).unsqueeze(0) < torch.randint(Lq // 2, Lq + 1, (B, 1), generator=gen)).to(dev)
w_c, w_q = (torch.randn(d, 1, generator=gen).to(dev) * 0.1 for _ in range(2))
w_cq = (torch.randn(1, 1, d, generator=gen) * 0.1).to(dev)
bias = torch.zeros(1, device=dev)
g = torch.randn(B, Lc, 4 * d, generator=gen).to(dev)


def fwd_bwd(fn, *leaves):
    ls = [t.detach().requires_grad_(True) for t in leaves]
    fn(*ls).backward(g)


for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    with torch.no_grad():
        t_f = timed(lambda: eager_bidaf(c, q, cm, qm, w_c, w_q, w_cq, bias))
    t_fb = timed(lambda: fwd_bwd(lambda cc, qq, a1, a2, a3: eager_bidaf(cc, qq, cm, qm, a1, a2, a3, bias), c, q, w_c, w_q, w_cq))
    print(f"torch eager BiDAF (cuBLAS {'TF32' if tf32 else 'fp32'} + ATen), config 2: forward {t_f:.0f} us, forward+backward {t_fb:.0f} us")
torch.backends.cuda.matmul.allow_tf32 = False
for prec, name in ((ops.PREC_FP32, "fp32 tier"), (ops.PREC_BF16, "bf16 tcgen05 tier")):
    with torch.no_grad():
        t_f = timed(lambda: F.bidaf_attention(c, q, cm, qm, w_c, w_q, w_cq, bias, None, None, 1.0, prec))
    t_fb = timed(lambda: fwd_bwd(lambda cc, qq, a1, a2, a3: F.bidaf_attention(cc, qq, cm, qm, a1, a2, a3, bias, None, None, 1.0, prec),
                                 c, q, w_c, w_q, w_cq))
    print(f"mmbidaf_b200 BiDAF, {name}, config 2: forward {t_f:.0f} us, forward+backward {t_fb:.0f} us (through autograd)")

# ---- bi-LSTM: cuDNN on a PackedSequence vs the persistent kernels ------------------------------------------------------
for L in (409, 1024):
    B, H = 32, 100
    x = torch.randn(B, L, H, generator=gen).to(dev)
    lengths = torch.randint(L // 2, L + 1, (B,), generator=gen)
    lengths[0] = L
    lstm = nn.LSTM(H, H, 1, batch_first=True, bidirectional=True).to(dev)

    def cudnn_fwd(xx=x):
        ls, idx = lengths.sort(0, descending=True)                 # encoding.py:91-101
        packed = pack_padded_sequence(xx[idx.to(dev)], ls, batch_first=True)
        out, _ = lstm(packed)
        out, _ = pad_packed_sequence(out, batch_first=True, total_length=L)
        return out[idx.argsort().to(dev)]

    def cudnn_fwd_bwd():
        xx = x.detach().requires_grad_(True)
        cudnn_fwd(xx).sum().backward()

    with torch.no_grad():
        t_f = timed(cudnn_fwd)
    t_fb = timed(cudnn_fwd_bwd)
    print(f"cuDNN nn.LSTM (packed, bidirectional, B={B}, L={L}, H={H}): forward {t_f:.0f} us = {t_f / L:.2f} us/step, "
          f"forward+backward {t_fb:.0f} us = {t_fb / L:.2f} us/step")
    len_d = lengths.to(torch.int32).to(dev)
    order = lengths.sort(0, descending=True)[1].to(torch.int32).to(dev)
    weights = [p.detach() for p in (lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, lstm.weight_ih_l0_reverse,
                                    lstm.weight_hh_l0_reverse, lstm.bias_ih_l0_reverse, lstm.bias_hh_l0_reverse)]

    def ours_fwd_bwd():
        xx = x.detach().requires_grad_(True)
        ws = [w.detach().requires_grad_(True) for w in weights]
        F.lstm_layer(xx, len_d, order, ws)[0].sum().backward()

    with torch.no_grad():
        t_f = timed(lambda: F.lstm_layer(x, len_d, order, weights))
    t_fb = timed(ours_fwd_bwd)
    print(f"mmbidaf_b200 bi-LSTM layer (input GEMM + persistent kernel), same shapes: forward {t_f:.0f} us = {t_f / L:.2f} us/step, "
          f"forward+backward {t_fb:.0f} us = {t_fb / L:.2f} us/step (incl. the weight-gradient GEMMs)")
