#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s33.log; : > $L
timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "model or plugin or classifier or graph" >> $L 2>&1
echo "pytest rc=$?" >> $L
timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
tail -8 $L
