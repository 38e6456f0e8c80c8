#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s5.log; : > $L
timeout -k 5 120 python tools/attn_bench.py >> $L 2>&1
timeout -k 5 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_s5_bench.json 2> gpurun_out/r2_s5_bench.err
echo "bench rc=$?" >> $L
tail -5 gpurun_out/r2_s5_bench.err >> $L
tail -12 $L
