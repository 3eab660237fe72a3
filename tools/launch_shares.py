"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows[1 + skip:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:60]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
unit = rows[1][iu]
print(f"{'kernel':60s} {'n':>5s} {'total_'+unit:>14s} {'avg_'+unit:>12s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:5d} {v[1]:14.0f} {v[1]/v[0]:12.0f} {100*v[1]/tot:6.1f}%")
