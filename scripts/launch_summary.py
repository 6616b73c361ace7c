"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            k = re.sub(r"\(.*", "", d["Kernel Name"])[:70]
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1e3 if d["Metric Unit"] in ("nsecond", "ns") else v * 1e3 if d["Metric Unit"] in ("msecond", "ms") else v
            agg.setdefault(k, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:70s} n={len(v):4d} avg={sum(v) / len(v):10.1f} us  share={100 * sum(v) / tot:5.1f}%")
