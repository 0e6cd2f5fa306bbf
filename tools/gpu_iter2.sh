#!/bin/bash
# iteration loop: kernel + step parity tests, the bench line, the per-kernel roofline table
mkdir -p gpurun_out
bash tools/gpu_check.sh > gpurun_out/check.txt 2>&1; grep -E "^==|FAILED|rel " gpurun_out/check.txt | head -40
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; cat gpurun_out/kernel_roofline.log | tail -n 25
