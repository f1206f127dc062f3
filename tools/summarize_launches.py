"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        rows.append((re.sub(r"\(.*", "", row["Kernel Name"]), v))
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot:.0f} us in total (cold-cache, serialised: compare shares)")
print(f"{'share':>6} {'total_us':>10} {'n':>5} {'avg_us':>8}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{100 * t / tot:5.1f}% {t:10.0f} {n:5d} {t / n:8.1f}  {k[:100]}")
