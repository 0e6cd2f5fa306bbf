#!/bin/bash
# ncu launch list of eager C2 steps (every kernel with its device time; cold-cache, serialised: compare shares)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph"
timeout 600 $B > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "== launches exit $?"; tail -2 gpurun_out/ncu1.log
python tools/launch_summary.py gpurun_out/launches.csv 40
