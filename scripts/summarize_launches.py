"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv, sys, collections, re
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
rows = rows[skip:]
agg = collections.OrderedDict()
for k, ns in rows:
    k = re.sub(r"\(.*", "", k)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(rows)} launches, {tot/1e6:.3f} ms total (cold-cache, serialised under ncu: compare shares)")
print(f"{'kernel':70s} {'count':>6s} {'ms':>9s} {'share':>7s} {'avg_us':>9s}")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {c:6d} {ns/1e6:9.3f} {100*ns/tot:6.1f}% {ns/c/1e3:9.1f}")
