"""Kernel-time breakdown of one training step with torch.profiler (cheap alternative to an ncu launch list).
    python tools/profile_step.py [--batch 32] [--top 40]
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmbidaf_b200.models import MMBiDAF  # noqa: E402
from mmbidaf_b200.synth import make_batch  # noqa: E402
from mmbidaf_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=bench.CFG3["batch"])
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--precision", default="fast")
args = ap.parse_args()
import mmbidaf_b200  # noqa: E402
mmbidaf_b200.set_precision(args.precision)
dev = torch.device("cuda:0")
torch.manual_seed(224)
model = MMBiDAF(bench.HIDDEN, bench.E_TEXT, bench.E_AUDIO, bench.E_IMAGE, dev, drop_prob=bench.DROP,
                max_transcript_length=bench.M).to(dev)
trainer = Trainer(model)
c = bench.CFG3
batch = make_batch(args.batch, c["lt"], c["la"], c["li"], c["t_dec"], seed=224).to(dev)
for _ in range(3):
    trainer.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        trainer.step(batch)
    torch.cuda.synchronize()
events = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
total = sum(e.device_time_total for e in events) / 2
print(f"CUDA kernel time per step: {total / 1e3:.2f} ms over {sum(e.count for e in events) // 2} launches")
for e in sorted(events, key=lambda e: -e.device_time_total)[:args.top]:
    print(f"{e.device_time_total / 2 / 1e3:8.3f} ms {e.count // 2:5d}x  {e.key[:120]}")
