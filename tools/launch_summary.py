#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> markdown on stdout.
usage: launch_summary.py <launches.csv> [title]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES, not absolutes).\n")
print(f"Total kernel time of the captured launches: {tot / 1e6:.2f} ms\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% |")
