"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
usage: python profiles/summarize_launches.py profiles/<file>.csv > profiles/<file>.summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] == "ns" else v
    name = r[ki].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={v[0]:5d} total_us={v[1]:11.1f} avg_us={v[1] / v[0]:9.2f} share={v[1] / tot:.3f}")
