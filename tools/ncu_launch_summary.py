#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (dev tool)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]
kn, mv, mu = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = re.sub(r"\(.*", "", r[kn])[:70]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k:72s} {cnt[k]:8d} {tot[k]:10.3f} {100 * tot[k] / total:6.2f}%")
