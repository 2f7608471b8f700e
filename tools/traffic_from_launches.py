"""profiles/traffic.json from the committed ncu launch lists (profiles/r02_launches_<mode>.csv): dram__bytes_read.sum +
dram__bytes_write.sum of all kernels of ONE interval (second repetition of a 3-interval clip, averaged over its intervals)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
detail = {}
for mode in ("dense", "dense_smooth", "dense_lowres", "block", "block_clip", "block_lowres", "linear", "linear_lowres"):
    p = os.path.join(ROOT, "profiles", f"r02_launches_{mode}.csv")
    if not os.path.exists(p):
        continue
    per = collections.OrderedDict()
    for r in csv.reader(open(p)):
        if len(r) > 14 and r[0].isdigit() and "at::" not in r[4]:    # not: torch kernels of the input generator
            per.setdefault(int(r[0]), {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
    items = list(per.values())
    half = items[len(items) // 2:]                         # second repetition: 3 intervals
    byt = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in half)
    ns = sum(v.get("gpu__time_duration.sum", 0) for v in half)
    out[mode] = byt / 3.0
    detail[mode] = {"kernels_per_interval": len(half) / 3.0, "ncu_duration_us_per_interval": ns / 3e3,
                    "dram_bytes_per_interval": byt / 3.0}
out["_detail"] = detail
out["_source"] = "profiles/r02_launches_<mode>.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; durations are cold-cache and serialised)"
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(detail, indent=1))
