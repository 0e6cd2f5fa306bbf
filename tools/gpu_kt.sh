#!/bin/bash
# kernel tests + step tests, attention timing, bench line (+ per-shape GEMM table), per-kernel roofline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/kernels.log 2>&1
echo "== kernels exit $? : $(tail -n 1 gpurun_out/kernels.log)"; grep -E "Error|FAILED|^E " gpurun_out/kernels.log | head -20
timeout 300 python tools/attn_time.py 2>&1 | tail -n 2
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_inference.py tests/test_gpu_finetune.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/step.log 2>&1
echo "== step exit $? : $(tail -n 1 gpurun_out/step.log)"; grep -E "Error|FAILED|^E " gpurun_out/step.log | head -20
timeout 900 python bench.py --no-cpu-baseline --gemm-table gpurun_out/gemm_table.md > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err; cat gpurun_out/gemm_table.md
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; grep -E "RoPE|qkv" gpurun_out/kernel_roofline.log | cut -c1-200
