#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh > gpurun_out/check.txt 2>&1; grep -E "^==|FAILED|rel " gpurun_out/check.txt | head -40
timeout 600 python -m pytest tests/test_gpu_vocab_parallel.py tests/test_gpu_inference.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/vp.log 2>&1
echo "== vocab_parallel + inference exit $? : $(tail -n 1 gpurun_out/vp.log)"; grep -E "^rank|Error|rel vs|FAILED" gpurun_out/vp.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; grep -E "scatter|adamw" gpurun_out/kernel_roofline.log
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 $B > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "== launches exit $?"
