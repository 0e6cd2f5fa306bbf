#!/bin/bash
# round-end evidence, part B: --set full captures (with source pages) of the attention kernels and of the GEMM family inside an
# eager C2 step, after a plain run of the same command (the launch list of the same command: tools/gpu_launch_list.sh)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph"
timeout 400 $B > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_pair|attn_dq_pair|attn_dkv_pair" -s 3 -c 3 -f -o gpurun_out/prof_attn_r02 $B > gpurun_out/ncu2.log 2>&1
echo "== prof_attn exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel" -s 6 -c 12 -f -o gpurun_out/prof_gemm_r02 $B > gpurun_out/ncu3.log 2>&1
echo "== prof_gemm exit $?"
for n in prof_attn_r02 prof_gemm_r02; do ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/$n.raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/$n.raw.csv; done
