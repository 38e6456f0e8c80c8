#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s32.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529"
timeout -k 5 600 $TR tools/config_bench.py cfg4_rows > gpurun_out/r2_s32_configs.jsonl 2>> $L
echo "configs rc=$?" >> $L
grep "rc=" $L; cut -c1-900 gpurun_out/r2_s32_configs.jsonl
