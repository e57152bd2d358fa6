"""Timeline of ONE graph-replayed training step (BASELINE config 3) from torch.profiler's kernel records: where the
5.9 ms go when independent branches overlap on side streams -- the serial chains, the idle gaps, what runs beside what.
    python tools/step_timeline.py [--min-us 15] [--csv gpurun_out/step_timeline.csv]
Kernel records under a profiler are slightly dilated; compare shares and gaps, not absolutes (bench.py has the step time).
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mmbidaf_b200  # noqa: E402
from mmbidaf_b200.models import MMBiDAF  # noqa: E402
from mmbidaf_b200.synth import make_batch  # noqa: E402
from mmbidaf_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--min-us", type=float, default=15.0, help="list kernels at least this long individually")
ap.add_argument("--csv", default="")
ap.add_argument("--precision", default="fast")
args = ap.parse_args()
mmbidaf_b200.set_precision(args.precision)
dev = torch.device("cuda:0")
torch.manual_seed(224)
model = MMBiDAF(bench.HIDDEN, bench.E_TEXT, bench.E_AUDIO, bench.E_IMAGE, dev, drop_prob=bench.DROP,
                max_transcript_length=bench.M).to(dev)
trainer = Trainer(model)
c = bench.CFG3
batch = make_batch(c["batch"], c["lt"], c["la"], c["li"], c["t_dec"], seed=224).to(dev)
trainer.capture(batch)
for _ in range(3):
    trainer.step_graphed()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer.step_graphed()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
rows = [(e.time_range.start - t0, e.time_range.end - t0, e.name) for e in ev]
end = max(r[1] for r in rows)
print(f"{len(rows)} kernel / memcpy records, span {end / 1e3:.3f} ms, sum of durations {sum(r[1] - r[0] for r in rows) / 1e3:.3f} ms")
if args.csv:
    with open(args.csv, "w") as f:
        f.write("start_us,end_us,name\n")
        for s, e, n in rows:
            f.write(f"{s:.2f},{e:.2f},\"{n[:100]}\"\n")


def short(n):
    n = n.replace("mmb::<unnamed>::", "").replace("void ", "")
    return n[:70]


# idle gaps (no kernel running at all) and concurrency profile
busy_until, idle = 0.0, 0.0
gaps = []
for s, e, n in rows:
    if s > busy_until:
        idle += s - busy_until
        if s - busy_until >= 3.0:
            gaps.append((busy_until, s))
    busy_until = max(busy_until, e)
print(f"idle (nothing running): {idle:.0f} us in total; gaps >= 3 us: {len(gaps)}, {sum(b - a for a, b in gaps):.0f} us")
# phases: split the step at the long serial kernels
print(f"\nkernels >= {args.min_us} us, in start order (start, duration, what else is running at its start):")
for i, (s, e, n) in enumerate(rows):
    if e - s >= args.min_us:
        beside = sum(1 for s2, e2, _ in rows if s2 <= s < e2) - 1
        print(f"  {s:9.1f} {e - s:8.1f}  +{beside}  {short(n)}")
# per 250-us bin: busy fraction of a single 'lane' and the top kernel
print("\nper 250 us: number of launches, union-busy us, top kernel by time")
nb = int(end // 250) + 1
for b in range(nb):
    lo, hi = b * 250.0, (b + 1) * 250.0
    seg = [(max(s, lo), min(e, hi), n) for s, e, n in rows if e > lo and s < hi]
    seg.sort()
    u, cur = 0.0, lo
    by = {}
    for s, e, n in seg:
        by[short(n)] = by.get(short(n), 0.0) + (e - s)
        if e > cur:
            u += e - max(s, cur)
            cur = e
    top = max(by.items(), key=lambda kv: kv[1]) if by else ("-", 0.0)
    print(f"  {lo:7.0f} {len(seg):4d} {u:6.0f}  {top[0]} ({top[1]:.0f} us)")
