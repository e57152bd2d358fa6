"""Print the clock64() phase stamps of CTA (0,0) of the three tensor-core backward kernels (debugging aid).
    MMB_BIDAF_BWD_TRACE=1 python tools/bidaf_bwd_trace.py
Stamps per kernel: start, X landed, then per tile [stage landed, first MMAs issued, first MMAs done, tile stored + sync],
then [last MMAs done, epilogue done].
"""
import os
import sys

os.environ["MMB_BIDAF_BWD_TRACE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

B, Lc, Lq, d = 64, 512, 256, 200
dev = "cuda"
gen = torch.Generator().manual_seed(224)
c = torch.randn(B, Lc, d, generator=gen).to(dev)
q = torch.randn(B, Lq, d, generator=gen).to(dev)
cm = torch.ones(B, Lc, dtype=torch.bool, device=dev)
qm = torch.ones(B, Lq, dtype=torch.bool, device=dev)
w = [torch.randn(d, generator=gen).to(dev) * 0.1 for _ in range(3)]
bias = torch.zeros(1, device=dev)
out, q2c, lr, lc_, bm, ws = ops.bidaf_fwd(c, q, cm, qm, w[0], w[1], w[2], bias, None, None, 1.0, 1, save=True)
G = torch.randn_like(out)
for _ in range(3):
    ops.bidaf_bwd(G, c, q, cm, qm, w[0], w[1], w[2], bias, None, None, 1.0, out, bm, q2c, lr, lc_, ws, 1)
torch.cuda.synchronize()
tr = ops.bidaf_bwd.last_trace.cpu()
for k, name in enumerate(["PT", "DC", "DQ"]):
    n = int(tr[k, 255])
    t = tr[k, :n] - tr[k, 0]
    print(f"{name}: {n} stamps, total {int(t[-1])} cycles")
    print("  prologue (X landed):", int(t[1]))
    body = t[2:n - 2].view(-1, 4)
    prev = t[1]
    for i, row in enumerate(body):
        a, b_, c_, d_ = (int(v) for v in row)
        print(f"  tile {i:2d}: wait stage {a - int(prev):6d}  issue {b_ - a:5d}  wait mma {c_ - b_:6d}  elementwise+sync {d_ - c_:6d}")
        prev = d_
    print(f"  last mma wait {int(t[n - 2]) - int(prev)}  epilogue {int(t[n - 1] - t[n - 2])}")
