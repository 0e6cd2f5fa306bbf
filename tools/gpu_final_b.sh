#!/bin/bash
# round-end evidence, part B: ncu launch list of eager C2 steps, then --set full captures (source pages) of the attention
# kernels and of the fused-epilogue GEMMs; each ncu run follows a plain run of the same command
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph"
timeout 400 $B > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "== launches exit $?"
python tools/launch_summary.py gpurun_out/launches.csv 36 > gpurun_out/launches_summary.md; head -30 gpurun_out/launches_summary.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_pair|attn_dq_pair|attn_dkv_pair" -s 3 -c 3 -f -o gpurun_out/prof_attn_r02 $B > gpurun_out/ncu2.log 2>&1
echo "== prof_attn exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel" -s 6 -c 10 -f -o gpurun_out/prof_gemm_r02 $B > gpurun_out/ncu3.log 2>&1
echo "== prof_gemm exit $?"
for n in prof_attn_r02 prof_gemm_r02; do ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/$n.raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/$n.raw.csv; done
