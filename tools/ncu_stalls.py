"""Warp-stall samples of one ncu capture (--import-source on), summed over the kernel and by opcode.
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep
"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
idx = {k: i for i, k in enumerate(hdr)}
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot, by, bywhat = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in data:
    toks = [t for t in r[idx["Source"]].split() if not t.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    by[op] += int(r[idx["# Samples"]] or 0)
    for k in st:
        v = int(r[idx[k]] or 0)
        tot[k] += v
        bywhat[op][k] += v
S = sum(tot.values())
print("samples", S)
for k, v in tot.most_common(10):
    print(f"  {k:26s} {v:8d} {100 * v / S:5.1f}%")
for op, v in by.most_common(12):
    print(f"  {op:10s} {v:7d} {100 * v / S:5.1f}%  ", ", ".join(f"{k[6:]}:{c}" for k, c in bywhat[op].most_common(4)))
