#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s30.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525"
timeout -k 5 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_s30_bench_n8.json 2> gpurun_out/r2_s30_bench_n8.err
echo "bench n8 rc=$?" >> $L
tail -3 gpurun_out/r2_s30_bench_n8.err >> $L
timeout -k 5 420 $TR tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
grep "rc=" $L; cut -c1-330 gpurun_out/r2_s30_bench_n8.json
