#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s11.log; : > $L
timeout -k 5 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_s11_pytest.log 2>&1
echo "pytest rc=$?" >> $L; tail -3 gpurun_out/r2_s11_pytest.log >> $L
timeout -k 5 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_s11_bench.json 2> gpurun_out/r2_s11_bench.err
echo "bench rc=$?" >> $L
timeout -k 5 300 python __graft_entry__.py smoke >> $L 2>&1
echo "smoke rc=$?" >> $L
tail -12 $L
