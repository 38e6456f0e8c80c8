#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s20.log; : > $L
timeout -k 5 900 python -m pytest tests -m gpu -x -q -k "plugin or emulated" >> $L 2>&1
echo "pytest-subset rc=$?" >> $L
timeout -k 5 900 python bench.py --no-gpu-reference > gpurun_out/r2_s20_bench.json 2> gpurun_out/r2_s20_bench.err
echo "bench rc=$?" >> $L
tail -5 gpurun_out/r2_s20_bench.err >> $L
python - >> $L <<'PY'
import json
d = json.loads(open("gpurun_out/r2_s20_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches")})
PY
tail -30 $L
