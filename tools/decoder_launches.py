"""Kernel sequence of one decoder step (forward and backward) inside a training step, from torch.profiler."""
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

mmbidaf_b200.set_precision("fast")
dev = torch.device("cuda:0")
torch.manual_seed(224)
model = MMBiDAF(bench.HIDDEN, bench.E_TEXT, bench.E_AUDIO, bench.E_IMAGE, dev, drop_prob=bench.DROP,
                max_transcript_length=bench.M).to(dev)
model.use_streams = False
trainer = Trainer(model)
c = bench.CFG3
batch = make_batch(c["batch"], c["lt"], c["la"], c["li"], c["t_dec"], seed=224).to(dev)
for _ in range(3):
    trainer.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer.step(batch)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
names = [(e.name, e.device_time) for e in ev]
for marker, title in (("dec_attn_partial", "forward"), ("dec_out_softmax_bwd", "backward")):
    idx = [i for i, (n, _) in enumerate(names) if marker in n]
    a, b = idx[3], idx[4]
    print(f"--- {title} step: {b - a} launches, {sum(t for _, t in names[a:b]):.1f} us of kernel time")
    for n, t in names[a:b]:
        print(f"{t:7.1f}  {n[:110]}")
