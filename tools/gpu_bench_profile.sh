#!/bin/bash
# step tests, the default bench line, then ncu: launch list + full captures of the hot kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -q --tb=short -p no:cacheprovider -m gpu -s > gpurun_out/step.log 2>&1
echo "== step exit $? : $(tail -n 1 gpurun_out/step.log)"
grep -E "grad rel|grad leaves|stage |losses" gpurun_out/step.log | head -40
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 $B > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "== launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|attn_fwd_kernel" -s 6 -c 8 -o gpurun_out/prof_fwd $B > gpurun_out/ncu2.log 2>&1
echo "== prof_fwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_dq_kernel|attn_dkv_kernel|segment_reduce_kernel|item_gather_kernel" -c 6 -o gpurun_out/prof_bwd $B > gpurun_out/ncu3.log 2>&1
echo "== prof_bwd exit $?"
ls -la gpurun_out/
