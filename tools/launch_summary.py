"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py --steps 1 --warmup W --no-graph
--sections step`: the LAST step's launches, grouped by kernel, as a markdown table.
    python tools/launch_summary.py gpurun_out/launches_step_final5.csv [--steps-in-run 4] [--top 40]
"""
import argparse
import csv
import collections

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--steps-in-run", type=int, default=4, help="warm-up + timed steps in the profiled run")
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
rows = []
with open(a.csv) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))
# the run = model/optimizer set-up launches + steps-in-run identical steps: take the last 1/steps of the launches after
# aligning on the step's last kernel (the fused optimizer update)
ends = [i for i, (n, _) in enumerate(rows) if "adadelta_clip_kernel" in n]
if len(ends) >= 2:
    step = rows[ends[-2] + 1:ends[-1] + 1]
else:
    step = rows[-(len(rows) // a.steps_in_run):]
total = sum(t for _, t in step)
own = [(n, t) for n, t in step if "mmb::" in n]
print(f"last step: {len(step)} launches, {total / 1e3:.2f} ms of kernel time; own kernels (`mmb::*`): {len(own)} launches, "
      f"{sum(t for _, t in own):.0f} us = {100 * sum(t for _, t in own) / total:.1f} %\n")
by = collections.OrderedDict()
for n, t in step:
    c = by.setdefault(n, [0, 0.0])
    c[0] += 1
    c[1] += t
print("| share | us | launches | kernel |\n|---|---|---|---|")
for n, (k, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"| {100 * t / total:.1f}% | {t:.0f} | {k} | `{n[:110]}` |")
