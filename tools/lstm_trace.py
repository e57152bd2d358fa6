"""Phase timeline of the persistent LSTM forward kernel (MMB_LSTM_TRACE): clock64 stamps of lane 0 of every warp of CTA (0, 0) for
steps 64..95 -- loop top, end of the FMA phase, activations done, h computed, before the barrier.
    python tools/lstm_trace.py [--len 409]
"""
import argparse
import os
import sys

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--len", type=int, default=409)
ap.add_argument("--hidden", type=int, default=100)
a = ap.parse_args()
NS, NW, NP = 32, 8, 5
trace = torch.zeros(NS * NW * NP, dtype=torch.int64, device="cuda")
os.environ["MMB_LSTM_TRACE"] = str(trace.data_ptr())
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

B, L, H = a.batch, a.len, a.hidden
gen = torch.Generator().manual_seed(0)
lengths = torch.full((B,), L, dtype=torch.int32)
len_d = lengths.cuda()
w_hh = ((torch.rand(2, 4 * H, H, generator=gen) - 0.5) * 0.2).cuda()
gx = torch.randn(B, L, 2, 4 * H, generator=gen).cuda()
for _ in range(2):
    ops.lstm_layer_fwd(gx.clone(), w_hh, len_d, None, B, L, H, 2, True)
torch.cuda.synchronize()
t = trace.cpu().view(NS, NW, NP)
nw = (2 * H + 31) // 32
names = ["top->fma_end", "fma_end->act", "act->h", "h->pre_bar", "pre_bar->next_top"]
print(f"B={B} L={L} H={H}: per-step cycles, median over steps 64..94, per warp (lane 0)")
for w in range(nw):
    d = [(t[:, w, i + 1] - t[:, w, i])[:-1] for i in range(4)] + [(t[1:, w, 0] - t[:-1, w, 4])]
    step = t[1:, w, 0] - t[:-1, w, 0]
    print(f"  warp {w}: step {int(step.median())}  " + "  ".join(f"{n} {int(x.median())}" for n, x in zip(names, d)))
t0 = t[:, :nw, 0].min(dim=1).values
print("  first-warp loop top to loop top:", int((t0[1:] - t0[:-1]).median()))
print("  spread of loop tops across warps (max - min), median:", int((t[:, :nw, 0].max(dim=1).values - t0).median()))
print("  spread of pre_bar across warps, median:", int((t[:, :nw, 4].max(dim=1).values - t[:, :nw, 4].min(dim=1).values).median()))
