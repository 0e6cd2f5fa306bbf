#!/bin/bash
# ncu --set full of ONE forward attention launch (dense rows) after a plain run of the same command
mkdir -p gpurun_out
export ATTN_CASE=${ATTN_CASE:-dense}
C="python tools/attn_time.py --fwd-only"
timeout 300 $C > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pair -s 3 -c 1 -f -o gpurun_out/prof_fwd_r2 $C > gpurun_out/ncu_fwd.log 2>&1
echo "== ncu exit $?"; tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu_fwd.log
