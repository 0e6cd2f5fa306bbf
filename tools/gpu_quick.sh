#!/bin/bash
# quick targeted run: pytest -k "$1" on the kernel/step tests, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -k "$1" -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/quick.log 2>&1
echo "== quick ($1) exit $? : $(tail -n 1 gpurun_out/quick.log)"; grep -E "^\[|Error|FAILED|^E " gpurun_out/quick.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
