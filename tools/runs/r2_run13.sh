#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s13.log; : > $L
timeout -k 5 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_s13_pytest.log 2>&1
echo "pytest rc=$?" >> $L; tail -3 gpurun_out/r2_s13_pytest.log >> $L
grep -n "FAILED\|Error" gpurun_out/r2_s13_pytest.log | head -10 >> $L
timeout -k 5 600 python tools/stem_bench.py >> $L 2>&1
timeout -k 5 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_s13_bench.json 2> gpurun_out/r2_s13_bench.err
echo "bench rc=$?" >> $L
tail -16 $L
