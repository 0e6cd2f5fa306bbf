#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "attention or swiglu or rmsnorm or scatter" -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/attn.log 2>&1
echo "== attention/swiglu/rmsnorm/scatter kernels exit $? : $(tail -n 1 gpurun_out/attn.log)"; grep -E "^\[|rel |Error|FAILED" gpurun_out/attn.log | head -30
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_finetune.py tests/test_gpu_loader.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/step.log 2>&1
echo "== step+finetune+loader exit $? : $(tail -n 1 gpurun_out/step.log)"; grep -E "grad rel|worst|Error|FAILED|C2|^E " gpurun_out/step.log | head -50
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-330 gpurun_out/bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench.json; grep -o '"roofline": {[^}]*}' gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; grep -E "scatter_add_sorted C2|rmsnorm|w1\|w3 \(gate|catalog" gpurun_out/kernel_roofline.log
