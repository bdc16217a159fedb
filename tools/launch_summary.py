#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count,
total and mean duration, share.  usage: launch_summary.py launches.csv [last_n_launches]"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[1:] if r[ix["Metric Name"]] == "gpu__time_duration.sum"]
if len(sys.argv) > 2:
    data = data[-int(sys.argv[2]):]
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"^void ", "", name)[:70]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"{len(data)} launches, {T:.3f} ms total (serialised, cold-cache ncu replay times)")
print(f"{'kernel':70s} {'n':>5s} {'total ms':>10s} {'mean us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:70s} {cnt[k]:5d} {v:10.3f} {1e3 * v / cnt[k]:10.1f} {100 * v / T:6.1f}%")
