#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s24.log; : > $L
for v in "" "fa:-DFUSED_SKIP_ATTN=1" "fe:-DFUSED_SKIP_EPI=1" "fs:-DFUSED_SKIP_STORE=1" "fae:-DFUSED_SKIP_ATTN=1 -DFUSED_SKIP_EPI=1"; do
  echo "== $v" >> $L
  MMPFN_VARIANT="$v" timeout -k 5 300 python tools/row_bench.py 2>&1 | grep "fused" >> $L
done
cat $L
