#!/usr/bin/env python
"""Instructions per bin by source region, from an ncu report captured with --import-source on.
    python profiles/ncu_regions.py REPORT.ncu-rep [bins]
Regions are given as (label, file, first_line_regex, last_line_regex) so they survive edits."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
B = float(sys.argv[2]) if len(sys.argv) > 2 else 524288.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fpath, hdr, acc = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":
        key = (fpath, int(r[0]))
        if key not in acc:
            acc[key] = (r[1].strip(), int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("# Samples")] or 0))
tot = sum(v[1] for v in acc.values())
print("total %.1f warp-instructions per bin" % (tot / B))
files = sorted(set(k[0] for k in acc))
for f in files:
    lines = sorted((l, v) for (ff, l), v in acc.items() if ff == f)
    # 12-line buckets
    bucket, cur = {}, None
    for l, (src, n, s) in lines:
        b = (l - 1) // 12
        bucket.setdefault(b, [0, 0, src])
        bucket[b][0] += n
        bucket[b][1] += s
    for b in sorted(bucket):
        n, s, src = bucket[b]
        if n / B >= 4:
            print("%-18s %4d-%4d %7.1f/bin  | %s" % (f, b * 12 + 1, b * 12 + 12, n / B, src[:70]))
