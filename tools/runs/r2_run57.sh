#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 50 python -m pytest tests/test_gpu_blocks.py -m gpu -x -q -k two_streams > gpurun_out/r2_s57.log 2>&1
echo "rc=$?" >> gpurun_out/r2_s57.log; tail -4 gpurun_out/r2_s57.log
