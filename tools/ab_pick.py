#!/usr/bin/env python
"""Reads tools/ab_bench.py's JSON lines and names the build variant with the best geometric-mean step time relative to
the default build (variants whose episode metrics differ from the first library's are disqualified)."""
import json
import math
import sys

score = {}
try:
    for ln in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ab.jsonl"):
        d = json.loads(ln)
        for wl, res in d.items():
            if wl == "bandwidth" or "default" not in res:
                continue
            for k, v in res.items():
                if not v["metrics_equal_to_first"]:
                    score[k] = float("inf")
                score[k] = score.get(k, 0.0) + math.log(v["ms_per_step"] / res["default"]["ms_per_step"])
except Exception:
    pass
print(min(score, key=score.get) if score else "full")
