#!/bin/bash
# 2 GPUs: dry run of the multi-GPU config script + N=2 bench (updated kernels)
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s9.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 5 600 $TR --master-port 29521 tools/config_bench.py dry >> $L 2>&1
echo "dry rc=$?" >> $L
timeout -k 5 600 python tools/config_bench.py cfg3 >> $L 2>&1
echo "cfg3 rc=$?" >> $L
timeout -k 5 900 $TR --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_s9_bench_n2.json 2> gpurun_out/r2_s9_bench_n2.err
echo "bench n2 rc=$?" >> $L
grep -v "^$" $L | grep -v "OMP_NUM\|\*\*\*\*" | tail -20
