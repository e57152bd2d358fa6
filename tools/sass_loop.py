"""Instruction mix between the last two BAR.SYNC of a kernel's SASS (the steady-state loop body)."""
import collections, re, subprocess, sys
lib, pattern = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, bodies = None, {}
for line in names.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); bodies[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        bodies[cur].append(line)
for fn, lines in bodies.items():
    if pattern not in fn:
        continue
    idx = [i for i, l in enumerate(lines) if "BAR.SYNC" in l]
    body = lines[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else lines
    cnt = collections.Counter()
    for l in body:
        m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            cnt[m.group(2).split(".")[0]] += 1
    print(fn[:90], "total", len(lines), "loop", len(body))
    print("  ", cnt.most_common(18))
