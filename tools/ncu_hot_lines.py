"""Top source lines of an ncu report by warp-stall samples (tools only).
usage: python tools/ncu_hot_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
if "Source" not in out.splitlines()[0] if out else True:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find the header row
hi = next(i for i, r in enumerate(rows) if any("Sampling" in c for c in r))
hdr = rows[hi]
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name:
            return i
    for i, h in enumerate(hdr):
        if name in h:
            return i
    return None
c_src = col("Source")
c_samp = col("# Samples") if col("# Samples") is not None else col("Warp Stall Sampling (All Samples)")
c_exec = col("Instructions Executed")
c_addr = col("Address")
tot = 0
agg = defaultdict(lambda: [0, 0])
for r in rows[hi + 1:]:
    if len(r) <= max(c_src, c_samp):
        continue
    try:
        s = int(float(r[c_samp] or 0))
    except ValueError:
        continue
    e = 0
    if c_exec is not None:
        try:
            e = int(float(r[c_exec] or 0))
        except ValueError:
            pass
    key = r[c_src].strip()[:150]
    agg[key][0] += s
    agg[key][1] += e
    tot += s
print("header:", [h for h in hdr][:12])
print(f"total samples {tot}")
for k, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * s / max(tot, 1):5.1f}%  samples {s:7d}  executed {e:10d}  {k}")
