#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 600 python tools/concurrency_probe.py > gpurun_out/r2_s41.log 2>&1
echo "rc=$?" >> gpurun_out/r2_s41.log
grep -v "Warning" gpurun_out/r2_s41.log | grep "rep \|tile " | cut -c1-330 | tail -24
