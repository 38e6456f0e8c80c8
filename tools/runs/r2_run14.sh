#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s14.log; : > $L
run() { echo "== $1" >> $L; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py >> $L 2>&1; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> $L 2>&1; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py 300 2000 1 >> $L 2>&1;  MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py 333 1999 0 8.0 >> $L 2>&1; }
run "x0:"
run "x1:-DATTN_ONE_ARRIVE=1"
cat $L
