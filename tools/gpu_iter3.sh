#!/bin/bash
# kernel + step parity tests (incl. world-size-2 vocab-parallel), bench line, roofline table
mkdir -p gpurun_out
bash tools/gpu_check.sh > gpurun_out/check.txt 2>&1; grep -E "^==|FAILED|rel " gpurun_out/check.txt | head -40
timeout 600 python -m pytest tests/test_gpu_vocab_parallel.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/vp.log 2>&1
echo "== vocab_parallel exit $? : $(tail -n 1 gpurun_out/vp.log)"; grep -E "rank|Error|error" gpurun_out/vp.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; cat gpurun_out/kernel_roofline.log | tail -n 25
