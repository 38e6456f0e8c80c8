#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s21.log; : > $L
run() { echo "== $1" >> $L; for a in "" "10000 10000 0 1.0 1 42" "300 2000 1" "333 1999 0 8.0" "2000 2000 0 1.0 4 20"; do MMPFN_VARIANT="$1" timeout -k 5 120 python tools/attn_bench.py $a >> $L 2>&1; done; }
run "np:-DATTN_PREFETCH=0"
run ""
run "np:-DATTN_PREFETCH=0"
run ""
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "attention or single_layer or large or eight" >> $L 2>&1
echo "pytest rc=$?" >> $L
timeout -k 5 300 python tools/step_bench.py 2>&1 | tail -1 >> $L
grep -v Warning $L | tail -60
