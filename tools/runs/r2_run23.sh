#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s23.log; : > $L
timeout -k 5 300 python -m pytest tests -m gpu -x -q -k "feature_qkv_attention or feature_attention_bf16" >> $L 2>&1
echo "unit rc=$?" >> $L
timeout -k 5 300 python tools/row_bench.py 2>&1 | grep "feature\|qkv_proj" >> $L
echo "row_bench rc=$?" >> $L
timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
MMPFN_DEBUG_LIB=1 MMPFN_FEAT_FUSED=0 timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
MMPFN_DEBUG_LIB=1 MMPFN_FEAT_FUSED=1 timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "single_layer or eight or multi_group or golden" >> $L 2>&1
echo "model rc=$?" >> $L
tail -40 $L
