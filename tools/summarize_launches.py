"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel and grid shape."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg, kind, tot = collections.OrderedDict(), collections.Counter(), 0.0
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dfb::", "")
    v = float(r["Metric Value"].replace(",", "")) / 1000.0
    tot += v
    kind[name] += v
    cnt[name] += 1
    a = agg.setdefault((name, r["Grid Size"]), [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"{len(rows)} launches, total {tot:.1f} us")
for k, v in kind.most_common():
    print(f"  {k:40s} n={cnt[k]:4d} {v:9.1f} us  {v / tot:6.1%}  avg {v / cnt[k]:6.1f}")
print()
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:36s} grid={k[1]:14s} n={a[0]:3d} total={a[1]:8.1f} us avg={a[1] / a[0]:6.1f}")
