"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/summarize_launches.py launches.csv [skip] [count] > profiles/xxx.md
"""
import csv, sys, re, collections
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}[unit]
    rows.append((r["Kernel Name"], ns, r.get("Grid Size", ""), r.get("Block Size", "")))
rows = rows[skip:skip + count]
agg = collections.OrderedDict()
for name, ns, g, b in rows:
    key = re.sub(r"\(.*", "", name)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"launches {len(rows)}  total device time {total/1e6:.3f} ms (ncu: cold-cache, serialised; compare shares)\n")
print("| kernel | launches | total ms | share |")
print("|---|---|---|---|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% |")
