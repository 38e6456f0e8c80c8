#!/bin/bash
# A/B of the L2-resident chunked schedule (tuning build): rounds of 147 tiles per chunk for the feature chain / out-proj->MLP
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s18.log; : > $L
export MMPFN_DEBUG_LIB=1
for cfg in "0 0" "2 3" "2 0" "0 3" "1 2" "3 4" "2 2" "4 6"; do
  set -- $cfg
  MMPFN_CHUNK_FEAT=$1 MMPFN_CHUNK_MLP=$2 timeout -k 5 200 python tools/step_bench.py 2>&1 | grep -v Warning | tail -1 >> $L
done
unset MMPFN_DEBUG_LIB
timeout -k 5 200 python tools/step_bench.py 2>&1 | tail -1 >> $L
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "chunked or multi_group or clf8 or layer" >> $L 2>&1
cat $L
