#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -k "step_tail" -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/tail.log 2>&1
echo "== step tail exit $? : $(tail -n 1 gpurun_out/tail.log)"; grep -E "Error|assert" gpurun_out/tail.log | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "== bench N=$N exit $?"; tail -n 1 gpurun_out/bench_$N.json | cut -c1-400; grep -o '"e2e": {[^}]*}' gpurun_out/bench_$N.json; tail -n 3 gpurun_out/bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_$N.json 2>> gpurun_out/bench_$N.err
echo "== bench reference N=$N exit $?"; cut -c1-200 gpurun_out/bench_ref_$N.json
