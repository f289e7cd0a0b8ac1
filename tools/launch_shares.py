"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares (markdown)."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
order = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    ms = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else (v if u in ("msecond", "ms") else v * 1e3))
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += ms
    order.append((name, ms))
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"launches {n}, total {tot:.1f} ms")
print("| share | total ms | launches | avg ms | kernel |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"| {100 * v[1] / tot:.1f}% | {v[1]:.2f} | {v[0]} | {v[1] / v[0]:.3f} | `{k[:110]}` |")
