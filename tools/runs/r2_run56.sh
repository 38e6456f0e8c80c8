#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s56.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 5 55 $TR tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
grep "rc=" $L; grep -o "mode=[a-z_]*: max |sharded - unsharded| logits = [0-9.e+-]*; repeat equal = [A-Za-z]*" $L | sort | uniq -c
