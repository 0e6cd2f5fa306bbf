#!/bin/bash
# N-GPU checks on one box: the data-parallel bench line at N ranks (and, for N=2, vocab-parallel parity over NCCL)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
if [ "$N" = "2" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/test_gpu_vocab_parallel.py > gpurun_out/vp_nccl_$N.log 2>&1
echo "== vp nccl N=$N exit $?"; grep -E "rank|Error" gpurun_out/vp_nccl_$N.log | head -12
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "== bench N=$N exit $?"; tail -n 1 gpurun_out/bench_$N.json | cut -c1-420; tail -n 3 gpurun_out/bench_$N.err
