#!/bin/bash
# 2 GPUs: sharded engine (estimator ownership + row sharding) against the unsharded engine; large-shape parity tests
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s15.log; : > $L
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
timeout -k 5 300 python tools/attn_bench.py >> $L 2>&1
echo "attn rc=$?" >> $L
timeout -k 5 600 python -m pytest tests -m gpu -x -q -k "blocks or large" >> $L 2>&1
echo "pytest rc=$?" >> $L
grep -v "^W1\|^\*\*\*" $L | tail -40
