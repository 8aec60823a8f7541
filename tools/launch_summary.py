"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total device time and share
per kernel.   python tools/launch_summary.py gpurun_out/launches.csv > profiles/rN_launches.txt"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = r[kn].split("(")[0][:70]
    agg[name][0] += 1
    agg[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms (ncu per-launch times: cold cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} launches={v[0]:4d} total_ms={v[1] / 1e6:10.3f} share={v[1] / tot:.4f}")
