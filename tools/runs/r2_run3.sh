#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
L=gpurun_out/r2_s3.log; : > $L
for pp in 4 8; do MMPFN_DEBUG_LIB=1 MMPFN_ATTN_PP=$pp timeout -k 5 120 python tools/attn_bench.py >> $L 2>&1; done
timeout -k 5 120 python tools/attn_bench.py >> $L 2>&1
timeout -k 5 120 python tools/attn_bench.py 300 2000 1 >> $L 2>&1
timeout -k 5 120 python tools/attn_bench.py 10000 10000 0 1.0 1 42 >> $L 2>&1
timeout -k 5 120 python tools/attn_trace.py >> $L 2>&1
timeout -k 5 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_s3_pytest.log 2>&1
echo "pytest rc=$?" >> $L; tail -3 gpurun_out/r2_s3_pytest.log >> $L
timeout -k 5 600 python tools/ref_probe.py > gpurun_out/r2_ref_probe.jsonl 2> gpurun_out/r2_ref_probe.err
echo "probe rc=$?" >> $L
timeout -k 5 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_s3_bench.json 2> gpurun_out/r2_s3_bench.err
echo "bench rc=$?" >> $L
timeout -k 5 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_s3_ref.json 2> gpurun_out/r2_s3_ref.err
echo "ref rc=$?" >> $L
tail -20 $L
