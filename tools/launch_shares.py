"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: python tools/launch_shares.py launches.csv [steps]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
agg = collections.OrderedDict()
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"| kernel | launches/step | us/step | share |\n|---|---|---|---|")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {k} | {c / steps:g} | {v / steps:.0f} | {100 * v / tot:.1f}% |")
print(f"| total | {sum(a[0] for a in agg.values()) / steps:g} | {tot / steps:.0f} | |")
