#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 50 python -m pytest tests/test_gpu_plugin.py -m gpu -x -q -k "replay_equals" > gpurun_out/r2_s58.log 2>&1
echo "rc=$?" >> gpurun_out/r2_s58.log; tail -5 gpurun_out/r2_s58.log
