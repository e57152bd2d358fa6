"""DRAM traffic per launch of the roofline kernels from an `ncu --set full` report -> profiles/<name>.json (read by bench.py's
`roofline.traffic`; never a literal in bench.py).
    ncu -i gpurun_out/r2_fwd_prof.ncu-rep --page raw --csv > gpurun_out/r2_fwd_raw.csv
    python tools/ncu_traffic.py gpurun_out/r2_fwd_raw.csv "bidaf_pack_kernel|bidaf_tc5_kernel" profiles/r02_bidaf_fwd_traffic.json
The capture must hold exactly one forward (or backward) call: the bytes of every matching launch are summed."""
import csv
import json
import re
import sys

src, pattern, dst = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"source": src, "pattern": pattern, "kernels": [], "dram_bytes_read": 0, "dram_bytes_write": 0, "time_us": 0.0}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if not re.search(pattern, name):
        continue
    rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    t = float(r[col["gpu__time_duration.sum"]])
    tu = units[col["gpu__time_duration.sum"]]
    t_us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(tu, 1.0)
    out["kernels"].append({"name": name[:80], "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "time_us": round(t_us, 2)})
    out["dram_bytes_read"] += int(rd)
    out["dram_bytes_write"] += int(wr)
    out["time_us"] += t_us
out["time_us"] = round(out["time_us"], 2)
out["note"] = "ncu --set full --cache-control none --clock-control none; per-launch times under ncu are serialised, not bench values"
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out)[:400])
