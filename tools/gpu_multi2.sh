#!/bin/bash
# 2 GPUs: NCCL vocab-parallel parity inside pytest, then the C2 / C4 / C4L bench lines at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/smi_multi.txt
timeout 600 python -m pytest tests/test_gpu_vocab_parallel.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/vp_nccl_pytest.log 2>&1
echo "== vp pytest exit $? : $(tail -n 1 gpurun_out/vp_nccl_pytest.log)"; grep -E "^E  |FAILED|rank " gpurun_out/vp_nccl_pytest.log | head
for c in C2 C4 C4L; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 6 --warmup 3 > gpurun_out/bench_${c}_n2.json 2> gpurun_out/bench_${c}_n2.err
  echo "== bench $c N=2 exit $?"; python - <<P
import json
try:
    j=json.load(open('gpurun_out/bench_${c}_n2.json'))
    print('$c N=2', round(j['ms_per_step'],2),'ms', round(j['value']), 'e2e', round(j['e2e']['value']), 'flags', j['status_flags'], j['config']['parallelism'])
except Exception as e: print('parse failed', e)
P
  tail -n 4 gpurun_out/bench_${c}_n2.err | cut -c1-300
done
