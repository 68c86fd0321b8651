#!/usr/bin/env python
"""Top stall-sample locations of one kernel from an ncu report:  python profiles/hotspots.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = {h: i for i, h in enumerate(rows[hdr])}
data = []
for k, r in enumerate(rows[hdr + 1:]):
    try:
        data.append((int(r[ci["# Samples"]]), k, r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f"total samples {tot}, instructions {len(data)}")
for s, k, src, ex in sorted(data, reverse=True)[:n]:
    print(f"{s:8d} {100 * s / tot:5.1f}%  #{k:5d} x{ex:9d}  {src[:100]}")
