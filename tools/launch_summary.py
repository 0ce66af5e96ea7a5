"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum). Usage: python tools/launch_summary.py launches.csv [reps]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
agg = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) < 15:
        continue
    agg.setdefault(r[4].split("(")[0], []).append(float(r[-1]) / 1e6)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"  {k:18s} n={len(v):4d} mean={sum(v)/len(v):8.3f} ms total={sum(v):9.2f} ms ({100*sum(v)/tot:4.1f}%)")
print(f"  all kernels {tot:.2f} ms")
