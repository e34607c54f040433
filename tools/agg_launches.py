"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the total)."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); n = 0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    agg[k][0] += 1; agg[k][1] += v; n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot:.1f} us total (per-launch times are cold-cache and serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={v[0]:5d} total={v[1]:10.1f}us mean={v[1]/v[0]:8.2f}us share={100*v[1]/tot:5.1f}%")
