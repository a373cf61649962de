"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid) -> markdown."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), ns))
tot = sum(r[3] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, g, b, ns in rows:
    k = (n, g) if "--by-grid" in sys.argv else (n,)
    agg[k][0] += 1
    agg[k][1] += ns
print(f"total launches {len(rows)}, total kernel time {tot/1e6:.2f} ms (serialised, cold-cache: compare SHARES)")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {' '.join(k)} | {c} | {ns/1e6:.3f} | {100*ns/tot:.1f}% | {ns/c/1e3:.1f} |")
