"""Warp-stall breakdown of one kernel from the source page of an `ncu --set full --import-source on` report:

    python tools/stall_summary.py <report.ncu-rep> [top-N instructions]

Prints the stall reasons over all samples, the samples per opcode with their dominant reasons, and the hottest
instructions.  (Warps that wait on an mbarrier show up as stall_long_sb on the BRA of the try_wait loop: producer
warps of a TMA pipeline spend most of their life there by design.)"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, byop, execd, per = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), []
samples = 0
for r in data:
    try:
        s = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    samples += s
    src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    op = (src.split()[0] if src else "?").split(".")[0]
    d = {h: int(r[ix[h]] or 0) for h in stalls}
    for h, v in d.items():
        tot[h] += v
        byop[op][h] += v
    execd[op] += int(r[ix["Instructions Executed"]] or 0)
    per.append((s, r[ix["Address"]][-6:], r[ix["Source"]].strip(), {h[6:]: v for h, v in d.items() if v}))
print("total samples", samples)
for h, v in tot.most_common():
    if v:
        print(f"  {h:28s} {v:8d} {100 * v / max(samples, 1):5.1f}%")
print("per opcode:")
for op, c in sorted(byop.items(), key=lambda x: -sum(x[1].values()))[:12]:
    s = sum(c.values())
    print(f"  {op:10s} samples {s:6d} {100 * s / max(samples, 1):5.1f}%  executed {execd[op]:10d}  ",
          [(k[6:], v) for k, v in c.most_common(4)])
print("hottest instructions:")
for s, a, src, d in sorted(per, key=lambda x: -x[0])[:top]:
    print(f"  {s:7d} {100 * s / max(samples, 1):5.2f}% {a} {src[:60]:60s} {sorted(d.items(), key=lambda x: -x[1])[:3]}")
