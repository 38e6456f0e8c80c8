#!/bin/bash
# 2 GPUs: sharded engine vs unsharded, then the N=2 bench line
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s4.log; : > $L
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dist.py >> $L 2>&1
echo "check_dist rc=$?" >> $L
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_s4_bench_n2.json 2> gpurun_out/r2_s4_bench_n2.err
echo "bench n2 rc=$?" >> $L
grep -c "AllGather" gpurun_out/r2_s4_bench_n2.err >> $L
tail -12 $L
