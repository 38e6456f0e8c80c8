#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s8.log; : > $L
run() { echo "== $1" >> $L; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py >> $L 2>&1; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> $L 2>&1; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py 300 2000 1 >> $L 2>&1; }
run "w0:"
run "w1:-DATTN_PP=4"
run "w2:-DATTN_PP=8"
run "w3:-DATTN_PP=7"
run "w4:-DATTN_PP=5"
cat $L
