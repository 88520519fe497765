"""Summarise an ncu gpu__time_duration launch list:  python scripts/launch_summary.py launches.csv"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:64]
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t:12.1f} us {c:5d} x {t / c:10.2f} us  {100 * t / tot:5.1f}%  {k}")
