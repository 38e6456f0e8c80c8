"""Aggregate the `ncu --page source --csv` dump of one kernel: stall reasons and samples per opcode.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_stalls.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = collections.Counter()
samples = 0
byop, execs = collections.Counter(), collections.Counter()
for r in data:
    n = int(r[ix["# Samples"]] or 0)
    samples += n
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    byop[op] += n
    execs[op] += int(r[ix["Instructions Executed"]] or 0)
print("total samples", samples, "instructions", len(data))
for s, v in tot.most_common():
    if v:
        print(f"{s:28s} {v:8d} {100 * v / samples:5.1f}%")
te = sum(execs.values())
print()
for op, v in byop.most_common(22):
    print(f"{op:12s} samples {v:7d} {100 * v / samples:5.1f}%   executed {execs[op]:12d} {100 * execs[op] / te:5.1f}%")
