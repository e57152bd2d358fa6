"""Print the clock64() phase stamps of CTA (0,0) of the two tensor-core forward kernels (debugging aid).
    python tools/bidaf_fwd_trace.py
Per tile: [stage landed, S MMAs issued, S MMAs done, row max exchanged, P stored + sync, (rescale) ready for P V].
"""
import os
import sys

os.environ["MMB_BIDAF_FWD_TRACE"] = "1"
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
for _ in range(3):
    ops.bidaf_fwd(c, q, cm, qm, w[0], w[1], w[2], bias, None, None, 1.0, 1)
torch.cuda.synchronize()
tr = ops.bidaf_fwd.last_trace.cpu()
for k, name in enumerate(["Q2C", "C2Q"]):
    n = int(tr[k, 255])
    t = tr[k, :n] - tr[k, 0]
    print(f"{name}: {n} stamps, total {int(t[-1])} cycles;  prologue (X landed) {int(t[1])}")
    nep = 4 if name == "Q2C" else 6          # stamps after the last P V wait: text tile landed, [start, drained] per accumulator, end
    last = n - 1 - nep
    ep = [int(v) for v in t[last:n]]
    body = t[2:last].view(-1, 6)
    prev = int(t[1])
    for i, row in enumerate(body):
        v = [int(x) for x in row]
        print(f"  tile {i}: wait stage {v[0] - prev:5d}  issue S {v[1] - v[0]:5d}  wait S {v[2] - v[1]:5d}  ld+max+exchange {v[3] - v[2]:5d}"
              f"  exp+P store+sync {v[4] - v[3]:5d}  rescale {v[5] - v[4]:5d}")
        prev = v[5]
    print(f"  last P V wait {ep[0] - prev}  epilogue steps {[b - a for a, b in zip(ep[:-1], ep[1:])]}")
