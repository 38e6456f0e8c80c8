#!/bin/bash
# 8 GPUs: sharded-vs-unsharded check, the N=8 bench line, configs[4] (256 packed tasks) and configs[3] (50k/50k) in their stated form
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s10.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -k 5 300 $TR --master-port 29531 tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
timeout -k 5 600 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_s10_bench_n8.json 2> gpurun_out/r2_s10_bench_n8.err
echo "bench n8 rc=$?" >> $L
timeout -k 5 400 $TR --master-port 29533 tools/config_bench.py cfg5 >> $L 2>&1
echo "cfg5 rc=$?" >> $L
timeout -k 5 700 $TR --master-port 29534 tools/config_bench.py cfg4 >> $L 2>&1
echo "cfg4 rc=$?" >> $L
grep -v "^$" $L | grep -v "OMP_NUM\|\*\*\*\*" | tail -30 | cut -c1-600
