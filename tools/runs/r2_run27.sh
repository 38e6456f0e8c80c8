#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s27.log; : > $L
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "feature_qkv or large or tasks or random" >> $L 2>&1
echo "pytest rc=$?" >> $L
timeout -k 5 600 python tools/config_bench.py cfg3 > gpurun_out/r2_s27_cfg3.jsonl 2>> $L
echo "cfg3 rc=$?" >> $L
tail -12 $L; cut -c1-700 gpurun_out/r2_s27_cfg3.jsonl
