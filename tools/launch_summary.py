"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
usage: python tools/launch_summary.py launches.csv [steps_in_capture] > summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
agg = collections.OrderedDict()
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"])
    n = re.sub(r"^void ", "", n).replace("mmpfn::<unnamed>::", "")[:60]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e6
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms of kernel time in the capture ({steps} step(s) + one-off fit/stage launches)")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    if a[1] / tot < 0.0005:
        continue
    print(f"{n:60s} n={a[0]:5d} total={a[1]:9.3f} ms share={100 * a[1] / tot:5.1f}% avg={1e3 * a[1] / a[0]:8.1f} us")
