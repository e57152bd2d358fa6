"""Stage-by-stage check of the fused BiDAF backward (bf16 tier) against fp64 torch on the GPU.
    python tools/bidaf_bwd_check.py [--shape B Lc Lq d] [--dropout]
Runs the launches cumulatively (MMB_BIDAF_BWD_STAGES) and prints max-norm relative errors of what each stage owns.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=4, default=[2, 130, 257, 200])
ap.add_argument("--dropout", action="store_true")
a = ap.parse_args()
B, Lc, Lq, d = a.shape
dev = "cuda"
gen = torch.Generator().manual_seed(7)
c = torch.randn(B, Lc, d, generator=gen).to(dev)
q = torch.randn(B, Lq, d, generator=gen).to(dev)
clen = torch.randint(1, Lc + 1, (B,), generator=gen); clen[0] = Lc
qlen = torch.randint(1, Lq + 1, (B,), generator=gen); qlen[0] = Lq
cm = (torch.arange(Lc).unsqueeze(0) < clen.unsqueeze(1)).to(dev)
qm = (torch.arange(Lq).unsqueeze(0) < qlen.unsqueeze(1)).to(dev)
w_c, w_q, w_x = (torch.randn(d, generator=gen).to(dev) * 0.1 for _ in range(3))
bias = torch.full((1,), 0.3, device=dev)
G = torch.randn(B, Lc, 4 * d, generator=gen).to(dev)
pr = 0.2 if a.dropout else 0.0
kc = (torch.rand(B, Lc, d, generator=gen) >= pr).to(dev) if a.dropout else None
kq = (torch.rand(B, Lq, d, generator=gen) >= pr).to(dev) if a.dropout else None
scale = 1.0 / (1.0 - pr)

out, q2c, lse_row, lse_col, bm, ws = ops.bidaf_fwd(c, q, cm, qm, w_c, w_q, w_x, bias, kc, kq, scale, 1, save=True)

# fp64 closed form
D = torch.float64
c64, q64, G64 = c.to(D), q.to(D), G.to(D)
cd = c64 if kc is None else c64 * kc * scale
qd = q64 if kq is None else q64 * kq * scale
S = (cd @ w_c.to(D)).unsqueeze(2) + (qd @ w_q.to(D)).unsqueeze(1) + (cd * w_x.to(D)) @ qd.transpose(1, 2) + bias.to(D)
neg = torch.full((), -1e30, dtype=D, device=dev)
P = torch.softmax(torch.where(qm.unsqueeze(1), S, neg), dim=2)
R = torch.softmax(torch.where(cm.unsqueeze(2), S, neg), dim=1)
A = P @ q64
T = R.transpose(1, 2) @ c64
Bm = P @ T
g0, g1, g2, g3 = G64.split(d, dim=2)
dA = g1 + c64 * g2
dBm = c64 * g3
dc0 = g0 + A * g2 + Bm * g3
dq1 = P.transpose(1, 2) @ dA
dT = P.transpose(1, 2) @ dBm
dP = dA @ q64.transpose(1, 2) + dBm @ T.transpose(1, 2)
dR = c64 @ dT.transpose(1, 2)
dS = P * (dP - (dP * P).sum(2, keepdim=True)) * qm.unsqueeze(1) + R * (dR - (dR * R).sum(1, keepdim=True)) * cm.unsqueeze(2)
rows, cols = dS.sum(2), dS.sum(1)
dsq = dS @ qd
dcd = rows.unsqueeze(2) * w_c.to(D) + dsq * w_x.to(D)
dqd = cols.unsqueeze(2) * w_q.to(D) + dS.transpose(1, 2) @ (cd * w_x.to(D))
kcs = 1.0 if kc is None else kc * scale
kqs = 1.0 if kq is None else kq * scale
dc_full = dc0 + R @ dT + dcd * kcs
dq_full = dq1 + dqd * kqs
dw_c = (cd * rows.unsqueeze(2)).sum((0, 1)); dw_q = (qd * cols.unsqueeze(2)).sum((0, 1)); dw_x = (cd * dsq).sum((0, 1))


def err(got, want):
    return float((got.to(D) - want).abs().max() / want.abs().max().clamp_min(1e-30))


print(f"fwd: out {err(out, torch.cat([c64, A, c64 * A, c64 * Bm], 2)):.2e}  bm {err(bm, Bm):.2e}  q2c {err(q2c, T):.2e}")
for stages, label in [(1, "prep"), (3, "prep+PT"), (7, "+DC"), (11, "prep+PT+DQ"), (31, "all")]:
    os.environ["MMB_BIDAF_BWD_STAGES"] = str(stages)
    d_text, d_mod, g_c, g_q, g_x, g_b = ops.bidaf_bwd(G, c, q, cm, qm, w_c, w_q, w_x, bias, kc, kq, scale, out, bm, q2c, lse_row,
                                                     lse_col, ws, 1)
    torch.cuda.synchronize()
    if stages == 1:
        print(f"{label}: d_text vs dc0 {err(d_text, dc0):.2e}")
    elif stages == 3:
        print(f"{label}: d_modality vs P^T dA {err(d_mod, dq1):.2e}")
    elif stages == 7:
        print(f"{label}: d_text {err(d_text, dc_full):.2e}   (R dT alone would be {err(d_text, dc0 + R @ dT):.2e}, dcd alone {err(d_text, dc0 + dcd * kcs):.2e})")
    elif stages == 11:
        print(f"{label}: d_modality {err(d_mod, dq_full):.2e}")
    else:
        print(f"{label}: d_text {err(d_text, dc_full):.2e} d_modality {err(d_mod, dq_full):.2e} dw_c {err(g_c, dw_c):.2e} "
              f"dw_q {err(g_q, dw_q):.2e} dw_x {err(g_x, dw_x):.2e} dbias {float(g_b):.3e} (exact {float(dS.sum()):.3e}, |dS| sum {float(dS.abs().sum()):.3e})")
