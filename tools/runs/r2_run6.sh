#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s6.log; : > $L
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-gpu-reference > gpurun_out/r2_s6_bench.json 2> gpurun_out/r2_s6_bench.err
echo "bench rc=$?" >> $L
CS=/usr/local/cuda/bin/compute-sanitizer
timeout -k 5 900 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_blocks.py -q -p no:cacheprovider -k "item_attention or fused_mlp or out_projection or item_qkv or feature_attention or linear_bf16 or layernorm" > gpurun_out/r2_sanitizer_memcheck_blocks.log 2>&1
echo "memcheck blocks rc=$?" >> $L; tail -4 gpurun_out/r2_sanitizer_memcheck_blocks.log >> $L
timeout -k 5 600 $CS --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?" >> $L; tail -4 gpurun_out/r2_sanitizer_memcheck_smoke.log >> $L
timeout -k 5 900 $CS --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_blocks.py -q -p no:cacheprovider -k "(item_attention and (128-48 or 5-49 or 1-1)) or (fused_mlp and (128 or 300)) or (out_projection and 300) or (item_qkv and 128) or (feature_attention and 7-3)" > gpurun_out/r2_sanitizer_racecheck_blocks.log 2>&1
echo "racecheck rc=$?" >> $L; tail -4 gpurun_out/r2_sanitizer_racecheck_blocks.log >> $L
timeout -k 5 600 $CS --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_blocks.py -q -p no:cacheprovider -k "(item_attention and 333) or (fused_mlp and 300) or (out_projection and 300) or (item_qkv and 333) or (feature_attention and 300)" > gpurun_out/r2_sanitizer_synccheck_blocks.log 2>&1
echo "synccheck rc=$?" >> $L; tail -4 gpurun_out/r2_sanitizer_synccheck_blocks.log >> $L
cat $L
