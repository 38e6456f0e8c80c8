#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s16.log; : > $L
timeout -k 5 300 python tools/overlap_probe.py >> $L 2>&1
echo "overlap rc=$?" >> $L
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s16_tests.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 gpurun_out/r2_s16_tests.log >> $L
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc=$?" >> $L
timeout -k 5 900 python bench.py > gpurun_out/r2_s16_bench.json 2> gpurun_out/r2_s16_bench.err
echo "bench rc=$?" >> $L
tail -30 $L
