#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s7.log; : > $L
run() { echo "== $1" >> $L; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py >> $L 2>&1; MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> $L 2>&1; }
run "v0:"
run "v1:-DATTN_PIN=1"
run "v2:-DATTN_PIN=1 -DATTN_PROBE_AT=16"
run "v3:-DATTN_PIN=1 -DATTN_PROBE_AT=20 -DATTN_KBEHIND=7"
run "v4:-DATTN_PROBE_AT=16"
run "v5:-DATTN_PIN=1 -DATTN_PROBE_AT=12 -DATTN_KAHEAD=6 -DATTN_KBEHIND=8"
run "v0:"
cat $L
