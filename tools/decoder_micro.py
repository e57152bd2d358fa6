"""Micro-benchmark of the fused decoder step (forward + backward) at config-3 shapes.
    python tools/decoder_micro.py [--batch 32] [--lt 409] [--steps 12]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200.layers import MultimodalAttentionDecoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--lt", type=int, default=409)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
B, Lt, H, E, M = a.batch, a.lt, 100, 300, 409
dev = "cuda"
torch.manual_seed(0)
dec = MultimodalAttentionDecoder(E, H, M).to(dev).train()
enc_a = torch.randn(B, Lt, 2 * H, device=dev, requires_grad=True)
enc_i = torch.randn(B, Lt, 2 * H, device=dev, requires_grad=True)
mask = torch.ones(B, M, dtype=torch.bool, device=dev)
sent = torch.randn(B, 1, E, device=dev)
for it in range(a.iters):
    ea, ei = enc_a * 1.0, enc_i * 1.0                         # fresh tensors -> fresh sequence state
    h = torch.randn(B, 1, H, device=dev, requires_grad=True)
    cell = torch.zeros(1, B, H, device=dev)
    cov = torch.zeros(B, Lt, 1, device=dev)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    loss = 0
    for s in range(a.steps):
        probs, h, cell, att, cov = dec(sent, h, cell, ea, ei, cov, mask)
        loss = loss - torch.log(probs[:, s] + 1e-12).sum() + torch.min(att, cov).sum()
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    print(f"iter {it}: fwd {e0.elapsed_time(e1) * 1e3 / a.steps:.1f} us/step, bwd {e1.elapsed_time(e2) * 1e3 / a.steps:.1f} us/step")
