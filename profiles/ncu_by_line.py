#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples from an ncu report (needs -lineinfo + --import-source).

    python profiles/ncu_by_line.py gpurun_out/prof.ncu-rep [kernel-index] [bins-per-launch]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bins = float(sys.argv[3]) if len(sys.argv) > 3 else 524288.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur, fpath = [], None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Function Name":
            cur = {"fn": r[1], "file": fpath, "lines": []}
            blocks.append(cur)
        elif r[0] == "Line No":
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur and r[0].isdigit() and r[2] == "-":
            h = cur["hdr"]
            cur["lines"].append((int(r[0]), r[1], int(r[h.index("Instructions Executed")] or 0),
                                 int(r[h.index("# Samples")] or 0), fpath))
    # group blocks by function (one function spans several files via inlining)
    fns = []
    for b in blocks:
        if not fns or fns[-1][0] != b["fn"] or (b["file"].endswith("bpp_step.cu") and any(f.endswith("bpp_step.cu") for f in fns[-1][2])):
            fns.append([b["fn"], [], []])
        fns[-1][1].extend(b["lines"])
        fns[-1][2].append(b["file"])
    fn = fns[which]
    tot = sum(l[2] for l in fn[1])
    samp = sum(l[3] for l in fn[1])
    print(fn[0][:100])
    print("total warp-instructions %d = %.1f per bin; %d stall samples" % (tot, tot / bins, samp))
    for ln, src, n, s, f in sorted(fn[1], key=lambda l: -l[2])[:45]:
        print("%6.1f/bin %5.1f%% inst %5.1f%% samples  %s:%d  %s" % (n / bins, 100.0 * n / max(tot, 1), 100.0 * s / max(samp, 1),
                                                                 f.split("/")[-1], ln, src.strip()[:90]))


if __name__ == "__main__":
    main()
