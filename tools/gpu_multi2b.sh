#!/bin/bash
# 2 GPUs, strict time limits: graft_allreduce test, C4 (vocab-parallel) bench line, C2 with --graft-comm
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_comm.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/comm_pytest.log 2>&1
echo "== comm pytest exit $? : $(tail -n 1 gpurun_out/comm_pytest.log)"; grep -E "^E  |FAILED" gpurun_out/comm_pytest.log | head
run() { # name, limit, args...
  local n=$1 lim=$2; shift 2
  timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 "$@" > gpurun_out/bench_${n}_n2.json 2> gpurun_out/bench_${n}_n2.err
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
run C4 240 --config C4
run C2gc 200 --config C2 --graft-comm
