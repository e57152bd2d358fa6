"""Micro-benchmark of the persistent LSTM kernels: one bidirectional layer, forward + backward.
    python tools/lstm_micro.py [--batch 32] [--len 409] [--hidden 100] [--iters 5]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--len", type=int, default=409)
ap.add_argument("--hidden", type=int, default=100)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
B, L, H = a.batch, a.len, a.hidden
dev = "cuda"
gen = torch.Generator().manual_seed(0)
lengths = torch.randint(L // 2, L + 1, (B,), generator=gen)
lengths[0] = L
order = torch.Tensor(lengths.tolist()).sort(0, descending=True)[1].to(torch.int32).to(dev)
len_d = lengths.to(torch.int32).to(dev)
w_hh = ((torch.rand(2, 4 * H, H, generator=gen) - 0.5) * 0.2).to(dev)
gx = torch.randn(B, L, 2, 4 * H, generator=gen).to(dev)
dout = torch.randn(B, L, 2 * H, generator=gen).to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
tf, tb = [], []
for it in range(a.iters):
    g = gx.clone()
    torch.cuda.synchronize()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    out, h_n, c_n, cell, _ = ops.lstm_layer_fwd(g, w_hh, len_d, order, B, L, H, 2, True)
    e1.record()
    ops.lstm_layer_bwd(g, cell, w_hh, len_d, order, dout, None, None, B, L, H, 2)
    e2.record()
    torch.cuda.synchronize()
    tf.append(e0.elapsed_time(e1) * 1e3)
    tb.append(e1.elapsed_time(e2) * 1e3)
print(f"B={B} L={L} H={H}: fwd {min(tf):.1f} us ({min(tf) / L:.3f} us/step)  bwd {min(tb):.1f} us ({min(tb) / L:.3f} us/step)")
