"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ix["Metric Unit"]]]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k[:80]}` | {n} | {v:.3f} | {v / n:.4f} | {100 * v / tot:.1f} % |")
print(f"| **total** | {sum(n for n, _ in agg.values())} | {tot:.3f} | | |")
