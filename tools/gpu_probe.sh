#!/bin/bash
# Run the GPU test groups in separate processes (a trapped kernel poisons its CUDA context), each
# under its own timeout, logging into gpurun_out/.  Usage: tools/gpu_probe.sh [group ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name timeout pytest-args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/probe_summary.log
  timeout -k 5 "$to" python -m pytest "$@" -q --no-header -p no:cacheprovider > "gpurun_out/probe_$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/probe_$name.log)" | tee -a gpurun_out/probe_summary.log
  grep -E "^(FAILED|ERROR)|Error|error:|assert|mbarrier timeout" "gpurun_out/probe_$name.log" | head -n 12 | tee -a gpurun_out/probe_summary.log
}
: > gpurun_out/probe_summary.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv | tee -a gpurun_out/probe_summary.log
groups=${@:-"f32blocks tcgemm layer_f32 layer_bf16 model_f32 model_bf16"}
for g in $groups; do
  case $g in
    f32blocks)  run f32blocks 300 tests/test_gpu_blocks.py -k "layernorm or linear_f32" ;;
    tcgemm)     run tcgemm 300 tests/test_gpu_blocks.py -k "linear_bf16" ;;
    layer_f32)  run layer_f32 600 tests/test_gpu_blocks.py -k "single_layer and fp32" ;;
    layer_bf16) run layer_bf16 600 tests/test_gpu_blocks.py -k "single_layer and bf16" ;;
    model_f32)  run model_f32 900 tests/test_gpu_model.py -k "fp32 or nan_column" ;;
    model_bf16) run model_bf16 900 tests/test_gpu_model.py -k "bf16" ;;
    all)        run all 1800 tests -m gpu ;;
  esac
done
echo "=== done" | tee -a gpurun_out/probe_summary.log
