#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 $B > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_pair|attn_dq_pair|attn_dkv_pair" -s 2 -c 1 -o gpurun_out/prof_attn_fwd $B > gpurun_out/ncu_a.log 2>&1
echo "== fwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_dq_pair|attn_dkv_pair" -s 2 -c 2 -o gpurun_out/prof_attn_bwd $B > gpurun_out/ncu_b.log 2>&1
echo "== bwd exit $?"
