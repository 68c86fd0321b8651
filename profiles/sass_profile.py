#!/usr/bin/env python
"""Dynamic SASS profile of one kernel from an ncu report (source page):
    python profiles/sass_profile.py rep.ncu-rep [rows]
prints warp-instructions executed per opcode class and (if `rows` is given) per row, plus the listing with
execution counts to stdout (one line per SASS instruction: index, executed, avg threads, samples, text)."""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = {h: i for i, h in enumerate(rows[hdr])}
nrows = int(sys.argv[2]) if len(sys.argv) > 2 else None
tot = 0
by = collections.Counter()
lines = []
for k, r in enumerate(rows[hdr + 1:]):
    try:
        ex = int(r[ci["Instructions Executed"]]); th = float(r[ci["Avg. Threads Executed"]] or 0); sm = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    src = r[ci["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    by[op.split(".")[0]] += ex
    tot += ex
    lines.append((k, ex, th, sm, src))
print(f"warp-instructions executed: {tot}" + (f" = {tot / nrows * 32:.1f} issue slots per 32 rows = {tot / nrows:.2f} per row" if nrows else ""))
for op, c in by.most_common(40):
    print(f"{op:12s} {c:12d} {100 * c / tot:5.1f}%" + (f"  {c / nrows:.3f}/row" if nrows else ""))
print()
for k, ex, th, sm, src in lines:
    print(f"#{k:5d} x{ex:9d} t{th:5.1f} s{sm:5d}  {src[:110]}")
