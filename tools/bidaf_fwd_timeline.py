"""Per-CTA timeline of the fused forward launch (debugging aid): start / loop end / end of every block by SM.
    python tools/bidaf_fwd_timeline.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import ops  # noqa: E402

B, Lc, Lq, d = 64, 512, 256, 200
dev = "cuda"
gen = torch.Generator().manual_seed(224)
c = torch.randn(B, Lc, d, generator=gen).to(dev)
q = torch.randn(B, Lq, d, generator=gen).to(dev)
clen = torch.randint(Lc // 2, Lc + 1, (B, 1), generator=gen)
qlen = torch.randint(Lq // 2, Lq + 1, (B, 1), generator=gen)
cm = (torch.arange(Lc).unsqueeze(0) < clen).to(dev)
qm = (torch.arange(Lq).unsqueeze(0) < qlen).to(dev)
w = [torch.randn(d, generator=gen).to(dev) * 0.1 for _ in range(3)]
bias = torch.zeros(1, device=dev)
nblk = B * (Lq // 128 + 2 * (Lc // 128))
times = torch.zeros(nblk, 12, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.bidaf_fwd(c, q, cm, qm, w[0], w[1], w[2], bias, None, None, 1.0, 1)
os.environ["MMB_BIDAF_FWD_CTA_TIMES"] = str(times.data_ptr())
ops.bidaf_fwd(c, q, cm, qm, w[0], w[1], w[2], bias, None, None, 1.0, 1)
torch.cuda.synchronize()
t = times.cpu()
t0 = int(t[:, 0].min())
nq = B * (Lq // 128)
nc = B * (Lc // 128)
print(f"launch span {int(t[:, 2].max()) - t0} ns over {nblk} blocks on {len(set(t[:, 3].tolist()))} SMs")
for name, sl in (("Q2C", slice(0, nq)), ("C2QA", slice(nq, nq + nc)), ("C2QB", slice(nq + nc, nblk))):
    x = t[sl]
    dur, loop, epi = x[:, 2] - x[:, 0], x[:, 1] - x[:, 0], x[:, 2] - x[:, 1]
    print(f"{name}: {x.shape[0]} blocks; start {int(x[:, 0].min()) - t0}..{int(x[:, 0].max()) - t0} ns; duration mean {dur.float().mean():.0f} "
          f"(min {int(dur.min())}, max {int(dur.max())}); wait+loop mean {loop.float().mean():.0f}; epilogue mean {epi.float().mean():.0f} "
          f"(min {int(epi.min())}, max {int(epi.max())})")
    if int(x[:, 4].max()) > 0:   # two-blocks-per-SM cut: finer stamps
        pro, xl, s0, nt = x[:, 4] - x[:, 0], x[:, 5] - x[:, 4], x[:, 6] - x[:, 5], x[:, 7].float()
        tl = (x[:, 1] - x[:, 6]).float()
        print(f"      prologue {pro.float().mean():.0f}; X tile wait {xl.float().mean():.0f}; first S tile {s0.float().mean():.0f}; "
              f"tiles {nt.mean():.1f}; rest of the loop {tl.mean():.0f} = {(tl / (nt - 1).clamp(min=1)).mean():.0f} ns per further tile")
        if int(x[:, 8].max()) > 0 and int(x[:, 11].max()) == 0:      # cut 3: [8] drained into staging, [9] plain text tile landed
            print("      epilogue: loop end -> drained %.0f; -> text tile %.0f; -> end %.0f" % (
                (x[:, 8] - x[:, 1]).float().mean(), (x[:, 9] - x[:, 8]).clamp(min=0).float().mean(),
                (x[:, 2] - torch.maximum(x[:, 8], x[:, 9])).float().mean()))
        elif int(x[:, 8].max()) > 0:
            e = x[:, 8:12] - torch.cat([x[:, 1:2], x[:, 8:11]], 1)
            print("      epilogue: wait text tile + drain rows 0-63 %.0f; store %.0f; drain rows 64-127 %.0f; store %.0f"
                  % tuple(e.float().mean(0).tolist()))
        first = x[:, 0] - t0 < 2000          # blocks of the first wave vs later ones
        for nm, m in (("first wave", first), ("later", ~first)):
            if int(m.sum()):
                print(f"      {nm}: {int(m.sum())} blocks; prologue {pro[m].float().mean():.0f}; X wait {xl[m].float().mean():.0f}; "
                      f"first S {s0[m].float().mean():.0f}; per further tile {(tl[m] / (nt[m] - 1).clamp(min=1)).mean():.0f}; "
                      f"epilogue {epi[m].float().mean():.0f}")
# occupancy over time: how many blocks are in their epilogue at each microsecond
span = int(t[:, 2].max()) - t0
for us in range(0, span // 1000 + 1, 4):
    now = t0 + us * 1000
    run = ((t[:, 0] <= now) & (t[:, 2] > now))
    in_epi = run & (t[:, 1] <= now)
    print(f"  t={us:3d} us: {int(run.sum()):3d} blocks resident ({int((run[:nq]).sum())} Q2C), {int(in_epi.sum()):3d} in their epilogue")
