#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s19.log; : > $L
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "qkv_scatter or emulated or single_layer or eight_estimators" >> $L 2>&1
echo "pytest-subset rc=$?" >> $L
timeout -k 5 300 python tools/row_bench.py >> $L 2>&1
echo "row_bench rc=$?" >> $L
timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s19_tests.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 gpurun_out/r2_s19_tests.log >> $L
cat $L | tail -40
