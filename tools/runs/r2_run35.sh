#!/bin/bash
# evidence for the final build: launch list of the step under ncu, full capture of the hot kernels
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s35.log; : > $L
timeout -k 5 300 python bench.py --profile --steps 1 --warmup 1 --no-graph >> $L 2>&1
echo "plain profile run rc=$?" >> $L
timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_v3.csv python bench.py --profile --steps 1 --warmup 1 --no-graph > gpurun_out/r2_s35_ncu_launch.log 2>&1
echo "launch list rc=$?" >> $L
timeout -k 5 120 python tools/prof_kernels.py >> $L 2>&1
echo "plain prof_kernels rc=$?" >> $L
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:"tc_|feat_qkv" --launch-skip 6 --launch-count 6 -o gpurun_out/prof_r02_v4_kernels -f python tools/prof_kernels.py > gpurun_out/r2_s35_ncu_full.log 2>&1
echo "ncu full rc=$?" >> $L
tail -8 $L
