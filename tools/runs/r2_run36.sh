#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s36.log; : > $L
for v in "noilp:-DFUSED_ILP=false" "" "noilp:-DFUSED_ILP=false" ""; do
  echo "== $v" >> $L
  MMPFN_VARIANT="$v" timeout -k 5 300 python tools/row_bench.py 2>&1 | grep "^S=" >> $L
done
timeout -k 5 300 python -m pytest tests -m gpu -x -q -k "feature_qkv or feature_attention" >> $L 2>&1
cat $L | tail -30
