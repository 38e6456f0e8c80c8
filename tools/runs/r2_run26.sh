#!/bin/bash
# 2 GPUs: sharded engine equality (estimator ownership and row sharding) and the bench line at N=2
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s26.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout -k 5 600 $TR tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
timeout -k 5 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_s26_bench_n2.json 2> gpurun_out/r2_s26_bench_n2.err
echo "bench n2 rc=$?" >> $L
tail -3 gpurun_out/r2_s26_bench_n2.err >> $L
grep -v "^W1\|^\*\*\*\|OMP_NUM" $L | cut -c1-260 | tail -20; tail -c 1500 gpurun_out/r2_s26_bench_n2.json
