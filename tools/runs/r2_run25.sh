#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s25.log; : > $L
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s25_tests.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 gpurun_out/r2_s25_tests.log >> $L
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc=$?" >> $L
timeout -k 5 300 python tools/row_bench.py 2>&1 | grep "S=2000\|S=2300" >> $L
timeout -k 5 900 python bench.py > gpurun_out/r2_s25_bench.json 2> gpurun_out/r2_s25_bench.err
echo "bench rc=$?" >> $L
tail -3 gpurun_out/r2_s25_bench.err >> $L
python - >> $L <<'PY'
import json
d = json.loads(open("gpurun_out/r2_s25_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
for k, v in d.get("kernels", {}).items():
    if isinstance(v, dict): print(k, round(v["ms"] * 1e3, 1), "us", round(v["frac_hbm"], 3), v.get("tflops"))
PY
tail -45 $L
