#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s28.log; : > $L
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s28_tests.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 gpurun_out/r2_s28_tests.log >> $L
timeout -k 5 600 python tools/config_bench.py cfg3 > gpurun_out/r2_s28_cfg3.jsonl 2>> $L
echo "cfg3 rc=$?" >> $L
timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
tail -12 $L; cut -c1-500 gpurun_out/r2_s28_cfg3.jsonl
