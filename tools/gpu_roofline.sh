#!/bin/bash
# per-kernel roofline table + one full ncu capture of the dominant GEMM launch (dram traffic)
mkdir -p gpurun_out
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; tail -n 30 gpurun_out/kernel_roofline.log
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel" -s 40 -c 6 -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1
echo "== prof_gemm exit $?"
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm.raw.csv 2>/dev/null
