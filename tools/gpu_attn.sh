#!/bin/bash
# attention iteration: kernel parity (softmax + SiLU), step-level parity, bench line, attention roofline rows
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "attention" -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/attn.log 2>&1
echo "== attention kernels exit $? : $(tail -n 1 gpurun_out/attn.log)"; grep -E "^\[|Error|FAILED|timeout|^E " gpurun_out/attn.log | head -30
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_inference.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/step.log 2>&1
echo "== step+inference exit $? : $(tail -n 1 gpurun_out/step.log)"; grep -E "Error|FAILED|^E " gpurun_out/step.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; grep -E "attention" gpurun_out/kernel_roofline.log
