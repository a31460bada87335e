"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, re, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; order = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    short = re.sub(r"\(.*", "", row["Kernel Name"]); short = re.sub(r"^void ", "", short).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    agg[short][0] += 1; agg[short][1] += v; tot += v
    order.append((short, v, row.get("Grid Size", ""), row.get("Block Size", "")))
print(f"total {tot/1e3:.2f} ms over {len(order)} launches (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t/1e3:8.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg={t/n:8.1f} us  {k[:80]}")
print("--- top individual launches")
for s, v, g, b in sorted(order, key=lambda x: -x[1])[:14]:
    print(f"{v:9.1f} us grid={g} block={b} {s[:60]}")
