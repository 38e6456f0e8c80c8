"""Per-kernel SASS opcode counts of libmmpfn_b200.so (the evidence that the tensor-core kernels are tcgen05 / TMEM /
TMA code): python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodalpfn_b200", "libmmpfn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "FFMA2", "FADD2", "HMMA", "LDSM", "LDGSTS", "FFMA"]
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").replace("mmpfn::", "")
        cur = re.sub(r"\(.*", "", cur)
        while cur in counts:
            cur += "'"
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        if op in KEYS:
            counts[cur][op] += 1
print(f"{os.path.basename(lib)}: {len(counts)} kernels, sm_100a SASS (cuobjdump -sass); instruction counts per kernel")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
tot = collections.Counter()
for k, c in counts.items():
    print(f"{k[:58]:58s} {total[k]:6d} " + " ".join(f"{c[x]:7d}" for x in KEYS))
    tot.update(c)
print(f"{'TOTAL':58s} {sum(total.values()):6d} " + " ".join(f"{tot[x]:7d}" for x in KEYS))
env = subprocess.run(["strings", lib], capture_output=True, text=True).stdout
print("environment variables named in the library:", sorted(set(re.findall(r"\bMMPFN_[A-Z_]+\b", env))) or "none")
