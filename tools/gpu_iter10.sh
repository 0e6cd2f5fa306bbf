#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer_surface.py tests/test_gpu_kernels.py -k "trainer or attention" -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/surface.log 2>&1
echo "== trainer surface + attention exit $? : $(tail -n 1 gpurun_out/surface.log)"; grep -E "Error|FAILED|^E " gpurun_out/surface.log | head -20
