"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg.setdefault(row["Kernel Name"][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):4d} avg={sum(v)/len(v):9.1f}us total={sum(v)/1e3:8.2f}ms {100*sum(v)/tot:5.1f}%")
