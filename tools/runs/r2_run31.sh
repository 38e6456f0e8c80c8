#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s31.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527"
timeout -k 5 600 $TR tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
grep "rows mode\|rc=" $L | cut -c1-200
