"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv, sys, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0]
    val = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += ns
total = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {t/1e6:10.3f} {t/n/1e3:10.1f} {100*t/total:6.1f}%")
print(f"{'TOTAL':60s} {sum(v[0] for v in tot.values()):8d} {total/1e6:10.3f}")
