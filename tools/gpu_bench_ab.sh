#!/bin/bash
# the bench line of this tree and of the round-1 tree (_r1 = git worktree of 196786e) back to back on the same box
mkdir -p gpurun_out
for i in 1 2; do
  (cd _r1 && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline) > gpurun_out/bench_r1_$i.json 2> gpurun_out/bench_r1_$i.err
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_new_$i.json 2> gpurun_out/bench_new_$i.err
done
python - <<'P'
import json
for n in ('r1_1','new_1','r1_2','new_2'):
    try:
        j=json.load(open(f'gpurun_out/bench_{n}.json')); print(n, round(j['ms_per_step'],2),'ms', round(j['value']),'e2e',round(j['e2e']['value']), j['clocks'])
    except Exception as e: print(n,'failed',e)
P
