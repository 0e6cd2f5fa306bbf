#!/bin/bash
# N-GPU checks: vocab-parallel parity over NCCL, then the data-parallel bench line at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/test_gpu_vocab_parallel.py > gpurun_out/vp_nccl_$N.log 2>&1
echo "== vp nccl N=$N exit $?"; grep -E "rank|Error" gpurun_out/vp_nccl_$N.log | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "== bench N=$N exit $?"; tail -n 1 gpurun_out/bench_$N.json; tail -n 3 gpurun_out/bench_$N.err
timeout 600 python -m pytest tests/test_gpu_vocab_parallel.py tests/test_gpu_kernels.py -k "vocab or scatter" -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/vp_gloo.log 2>&1
echo "== vp gloo + scatter exit $? : $(tail -n 1 gpurun_out/vp_gloo.log)"; grep -E "^rank|Error" gpurun_out/vp_gloo.log | head
