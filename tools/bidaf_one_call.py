"""ONE fused BiDAF forward (or forward + backward with --bwd) at BASELINE config 2 after warm-up: the program an `ncu --set full`
capture of the roofline kernels runs (-k regex:bidaf -s <launches of the warm-up> ...)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

B, Lc, Lq, d = 64, 512, 256, 200
dev = "cuda"
gen = torch.Generator().manual_seed(224)
c, q = torch.randn(B, Lc, d, generator=gen).to(dev), torch.randn(B, Lq, d, generator=gen).to(dev)
cm = (torch.arange(Lc).unsqueeze(0) < torch.randint(Lc // 2, Lc + 1, (B, 1), generator=gen)).to(dev)
qm = (torch.arange(Lq).unsqueeze(0) < torch.randint(Lq // 2, Lq + 1, (B, 1), generator=gen)).to(dev)
w = [torch.randn(d, generator=gen).to(dev) * 0.1 for _ in range(3)]
bias = torch.zeros(1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(3):
    flush.zero_()                                           # the inputs are cold in L2, as between two steps of a training run
    if "--bwd" not in sys.argv:                             # what bench.py's roofline section times: nothing saved for a backward pass
        ops.bidaf_fwd(c, q, cm, qm, w[0], w[1], w[2], bias, precision=ops.PREC_BF16, aux=False)
        continue
    out, q2c, lr, lc_, bm, ws = ops.bidaf_fwd(c, q, cm, qm, w[0], w[1], w[2], bias, precision=ops.PREC_BF16, save=True)
    if "--bwd" in sys.argv:
        g = torch.randn_like(out) if it == 0 else g
        flush.zero_()
        ops.bidaf_bwd(g, c, q, cm, qm, w[0], w[1], w[2], bias, None, None, 1.0, out, bm, q2c, lr, lc_, ws, ops.PREC_BF16)
torch.cuda.synchronize()
