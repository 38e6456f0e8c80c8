#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_s1.log
timeout -k 5 300 python -m pytest tests/test_gpu_blocks.py -q -x -k "item_attention" -p no:cacheprovider >> gpurun_out/r2_s1.log 2>&1
echo "rc=$?" >> gpurun_out/r2_s1.log
timeout -k 5 120 python tools/attn_bench.py >> gpurun_out/r2_s1.log 2>&1
timeout -k 5 120 python tools/attn_bench.py 300 2000 1 >> gpurun_out/r2_s1.log 2>&1
timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> gpurun_out/r2_s1.log 2>&1
for pp in 0 3 4 8 10 12; do
  MMPFN_DEBUG_LIB=1 MMPFN_ATTN_PP=$pp timeout -k 5 120 python tools/attn_bench.py >> gpurun_out/r2_s1.log 2>&1
done
MMPFN_DEBUG_LIB=1 MMPFN_ATTN_PP=8 timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> gpurun_out/r2_s1.log 2>&1
MMPFN_DEBUG_LIB=1 MMPFN_ATTN_PP=4 timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> gpurun_out/r2_s1.log 2>&1
timeout -k 5 600 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_model.py -q -p no:cacheprovider >> gpurun_out/r2_s1.log 2>&1
echo "rc=$?" >> gpurun_out/r2_s1.log
tail -5 gpurun_out/r2_s1.log
