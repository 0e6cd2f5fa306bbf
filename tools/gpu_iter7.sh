#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "lora or jagged or attention" -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/lora_k.log 2>&1
echo "== lora kernels exit $? : $(tail -n 1 gpurun_out/lora_k.log)"; grep -E "^\[|rel |Error|FAILED|assert" gpurun_out/lora_k.log | head -30
timeout 600 python -m pytest tests/test_gpu_finetune.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/finetune.log 2>&1
echo "== finetune exit $? : $(tail -n 1 gpurun_out/finetune.log)"; grep -E "grad rel|losses|worst|Error|FAILED" gpurun_out/finetune.log | head -60
timeout 900 python -m pytest tests/test_gpu_step.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/step.log 2>&1
echo "== step exit $? : $(tail -n 1 gpurun_out/step.log)"; grep -E "losses|worst|Error|FAILED|C2" gpurun_out/step.log | head -30
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-330 gpurun_out/bench.json; grep -o '"cuda_graph": [a-z]*' gpurun_out/bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
