#!/bin/bash
# Round-end evidence: parity suite, the bench line, launch list, full ncu captures of the dominant kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $? : $(tail -n 1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
echo "== bench reference exit $?"; cat gpurun_out/bench_reference.json
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph"
timeout 600 $B > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
echo "== launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel" -s 1 -c 4 -o gpurun_out/prof_gemm_trunk $B > gpurun_out/ncu_g.log 2>&1
echo "== prof_gemm_trunk exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_pair|attn_dq_pair|attn_dkv_pair|segment_reduce|item_gather" -c 6 -o gpurun_out/prof_attn $B > gpurun_out/ncu_a.log 2>&1
echo "== prof_attn exit $?"
for f in prof_gemm_trunk prof_attn; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out | head -40
