"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [first_id last_id] > profiles/x.md"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
data = [r for r in rows[h + 1:] if len(r) > vi]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = collections.OrderedDict()
tot = 0.0
for r in data:
    if not (lo <= int(r[idi]) <= hi):
        continue
    name = r[ki].split("(")[0].replace("void ", "")[:70]
    ns = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f} % |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot / 1e6:.3f} | |")
