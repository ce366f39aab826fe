#!/usr/bin/env python
"""tools/launch_summary.py LAUNCHES.csv -- per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
(profiles/rNN_launches*.csv): launches, total time, share.  The times are cold-cache and serialised: shares, not absolutes."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
hdr = rows[0]
name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= val_i:
        continue
    v = float(r[val_i].replace(",", ""))
    u = r[unit_i]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    nm = r[name_i].split("(")[0]
    tot[nm] += us
    cnt[nm] += 1
total = sum(tot.values())
print("# per-kernel totals of %s (%d launches; cold-cache, serialised: shares, not absolute times)" % (sys.argv[1], sum(cnt.values())))
for nm in sorted(tot, key=lambda k: -tot[k]):
    print("%-70s launches %4d  total %12.1f us  share %6.2f %%" % (nm[:70], cnt[nm], tot[nm], 100 * tot[nm] / total))
