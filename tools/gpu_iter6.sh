#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_finetune.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/finetune.log 2>&1
echo "== finetune exit $? : $(tail -n 1 gpurun_out/finetune.log)"; grep -E "losses|worst|rel|Error|FAILED|assert" gpurun_out/finetune.log | head -30
bash tools/gpu_check.sh > gpurun_out/check.txt 2>&1; grep -E "^==|FAILED|rel " gpurun_out/check.txt | head -40
timeout 600 python -m pytest tests/test_gpu_vocab_parallel.py tests/test_gpu_inference.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/vp.log 2>&1
echo "== vocab_parallel + inference exit $? : $(tail -n 1 gpurun_out/vp.log)"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-330 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
