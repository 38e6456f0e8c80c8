#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s54.log; : > $L
timeout -k 5 300 python tools/concurrency_probe.py 2>&1 | grep -v Warning | grep "pairs that differ\|'mlp'\|'out_proj_ln'\|rows that differ\|chain A alone" | cut -c1-160 >> $L
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s54_tests.log 2>&1
echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_s54_tests.log >> $L
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc=$?" >> $L
timeout -k 5 900 python bench.py --no-gpu-reference > gpurun_out/r2_s54_bench.json 2> gpurun_out/r2_s54_bench.err
echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d = json.loads(open("gpurun_out/r2_s54_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"])
for k in ("out_proj_residual_layernorm", "mlp_fused"):
    print(k, round(d["kernels"][k]["ms"] * 1e3, 1), "us")
PY
cat $L
