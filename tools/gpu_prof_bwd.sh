#!/bin/bash
# ncu --set full of one dQ and one dK/dV launch (dense rows) after a plain run of the same command
mkdir -p gpurun_out
export ATTN_CASE=${ATTN_CASE:-dense}
C="python tools/attn_time.py"
timeout 300 $C > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_dq_pair|attn_dkv_pair" -s 4 -c 2 -f -o gpurun_out/prof_bwd_r2 $C > gpurun_out/ncu_bwd.log 2>&1
echo "== ncu exit $?"; tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu_bwd.log
