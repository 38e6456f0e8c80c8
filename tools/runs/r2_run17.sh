#!/bin/bash
# 8 GPUs: row-sharded context build against the unsharded engine, and configs[3] with one / eight estimators
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s17.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout -k 5 420 $TR tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
timeout -k 5 600 $TR tools/config_bench.py cfg4_one_rows cfg4_one cfg4_rows > gpurun_out/r2_s17_configs.jsonl 2>> $L
echo "configs rc=$?" >> $L
grep -v "^W1\|^\*\*\*\|OMP_NUM" $L | tail -30; cat gpurun_out/r2_s17_configs.jsonl
