#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s2.log; : > $L
timeout -k 5 120 python tools/attn_trace.py >> $L 2>&1
MMPFN_ATTN_PP=0 timeout -k 5 120 python tools/attn_trace.py >> $L 2>&1
timeout -k 5 900 python -m pytest tests/test_gpu_large_parity.py -q -s -k "not ctx10k" -p no:cacheprovider >> $L 2>&1
echo "rc=$?" >> $L
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:tc_item_attn --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r02_v1_attn -f python tools/prof_attn.py > gpurun_out/r2_s2_ncu.log 2>&1
tail -3 gpurun_out/r2_s2_ncu.log >> $L
tail -30 $L
