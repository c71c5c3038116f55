"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, average, share."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for c in sys.argv[2:]:
    print("# " + c)
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:60]:60s} n={v[0]:5d} total_ms={v[1]/1e3:9.2f} avg_us={v[1]/v[0]:9.1f} share={v[1]/tot:.3f}")
