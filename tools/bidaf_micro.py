"""Micro-benchmark of the fused BiDAF forward at BASELINE config 2 (or --shape B Lc Lq), both tiers.
    python tools/bidaf_micro.py [--precision 0|1] [--iters 20]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", type=int, default=1)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--bwd", action="store_true", help="time the fused backward (bf16 tier) instead of the forward")
ap.add_argument("--dropout", action="store_true")
ap.add_argument("--shape", type=int, nargs=3, default=[64, 512, 256])
a = ap.parse_args()
B, Lc, Lq = a.shape
d = 200
dev = "cuda"
gen = torch.Generator().manual_seed(224)
sets = []
for _ in range(4):
    c = torch.randn(B, Lc, d, generator=gen).to(dev)
    q = torch.randn(B, Lq, d, generator=gen).to(dev)
    cm = (torch.arange(Lc).unsqueeze(0) < torch.randint(Lc // 2, Lc + 1, (B, 1), generator=gen)).to(dev)
    qm = (torch.arange(Lq).unsqueeze(0) < torch.randint(Lq // 2, Lq + 1, (B, 1), generator=gen)).to(dev)
    sets.append((c, q, cm, qm))
w = [torch.randn(d, generator=gen).to(dev) * 0.1 for _ in range(3)]
bias = torch.zeros(1, device=dev)
pr = 0.2 if a.dropout else 0.0
keeps = [((torch.rand(B, Lc, d, generator=gen) >= pr).to(dev), (torch.rand(B, Lq, d, generator=gen) >= pr).to(dev))
         if a.dropout else (None, None) for _ in range(4)]
if a.bwd:
    saved = []
    for s, k in zip(sets, keeps):
        out, q2c, lr, lc_, bm, ws = ops.bidaf_fwd(s[0], s[1], s[2], s[3], w[0], w[1], w[2], bias, k[0], k[1], 1 / (1 - pr),
                                                  a.precision, save=True)
        saved.append((torch.randn_like(out), out, q2c, lr, lc_, bm, ws))
    sets = [s + k + v for s, k, v in zip(sets, keeps, saved)]
    run = lambda s: ops.bidaf_bwd(s[6], s[0], s[1], s[2], s[3], w[0], w[1], w[2], bias, s[4], s[5], 1 / (1 - pr), s[7], s[11], s[8],
                                  s[9], s[10], s[12], a.precision)
else:
    sets = [s + k for s, k in zip(sets, keeps)]
    run = lambda s: ops.bidaf_fwd(s[0], s[1], s[2], s[3], w[0], w[1], w[2], bias, s[4], s[5], 1 / (1 - pr),
                                  precision=a.precision)
for i in range(5):
    run(sets[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.iters):
    run(sets[i % 4])
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / a.iters * 1e3
algo = 4 * B * (Lc * d + Lq * d + Lc * 4 * d) + B * (Lc + Lq)
if a.bwd:      # read dX (4 Lc d) + c, q; write dc, dq
    algo = 4 * B * (4 * Lc * d + 2 * Lc * d + 2 * Lq * d)
print(f"precision={a.precision} B={B} Lc={Lc} Lq={Lq}: {t:.1f} us/{'backward' if a.bwd else 'forward'}, {algo / t / 1e3:.1f} GB/s algorithmic "
      f"({algo / t / 1e3 / 6553.3 * 100:.1f}% of 6553 GB/s)")
