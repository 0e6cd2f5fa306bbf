#!/bin/bash
# Runs the GPU test groups in separate processes (a sticky CUDA error in one group
# must not hide the others), each under its own timeout; logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  timeout "$to" python -m pytest "$@" -q --tb=short -p no:cacheprovider -m gpu > "gpurun_out/$name.log" 2>&1
  echo "== $name exit $? : $(tail -n 1 gpurun_out/$name.log)"
}
run gemm_layouts 300 tests/test_gpu_kernels.py -k "gemm_layouts"
run gemm_other   300 tests/test_gpu_kernels.py -k "gemm and not layouts"
run attention    300 tests/test_gpu_kernels.py -k "attention"
run embed        300 tests/test_gpu_kernels.py -k "gather or scatter or jagged or action"
run elementwise  300 tests/test_gpu_kernels.py -k "rmsnorm or rope or swiglu or casts or rating"
run step         600 tests/test_gpu_step.py
grep -hE "^\[|rel |FAILED|PASSED|Error|error|losses|stage " gpurun_out/*.log | head -150
