#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_inference.py -k "attention or retrieval" -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/attention.log 2>&1
echo "== attention+retrieval tests exit $? : $(tail -n 1 gpurun_out/attention.log)"; grep -E "^E  |FAILED|^\[" gpurun_out/attention.log | head -12
bash tools/gpu_prof_fwd.sh
bash tools/gpu_prof_bwd.sh
