#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s29.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523"
timeout -k 5 600 $TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_s29_bench_n4.json 2> gpurun_out/r2_s29_bench_n4.err
echo "bench n4 rc=$?" >> $L
tail -3 gpurun_out/r2_s29_bench_n4.err >> $L
timeout -k 5 600 $TR bench.py --impl reference --gpus 4 --steps 1 --warmup 0 > gpurun_out/r2_s29_ref_n4.json 2>> $L
echo "ref arm n4 rc=$?" >> $L
cat $L | tail; cut -c1-400 gpurun_out/r2_s29_bench_n4.json; cut -c1-600 gpurun_out/r2_s29_ref_n4.json
