#!/bin/bash
# 4 GPUs, strict time limit: the C2 data-parallel bench line
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_n4.json 2> gpurun_out/bench_C2_n4.err
echo "== bench C2 N=4 exit $?"; python - <<P
import json
try:
    line=[l for l in open('gpurun_out/bench_C2_n4.json') if l.startswith('{')][-1]
    j=json.loads(line)
    print('C2 N=4', round(j['ms_per_step'],2),'ms', round(j['value']), 'e2e', round(j['e2e']['value']), 'flags', j['status_flags'], j['config']['parallelism'])
except Exception as e: print('parse failed', e)
P
tail -n 2 gpurun_out/bench_C2_n4.err | cut -c1-200
