#!/bin/bash
# 2 GPUs, strict time limits: C2 data-parallel, C4 and C4L (vocab-parallel head, 100 k / 1 M items) bench lines
mkdir -p gpurun_out
run() { # name, limit, args...
  local n=$1 lim=$2; shift 2
  timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_${n}_n2.json 2> gpurun_out/bench_${n}_n2.err
  echo "== bench $n N=2 exit $?"; python - <<P
import json
try:
    line=[l for l in open('gpurun_out/bench_${n}_n2.json') if l.startswith('{')][-1]
    j=json.loads(line)
    print('$n N=2', round(j['ms_per_step'],2),'ms', round(j['value']), 'e2e', round(j['e2e']['value']), 'flags', j['status_flags'], j['config']['parallelism'], j['config'].get('cuda_graph'))
except Exception as e: print('parse failed', e)
P
  tail -n 3 gpurun_out/bench_${n}_n2.err | cut -c1-300
}
run C2 200 --config C2
run C4 200 --config C4
run C4L 240 --config C4L
