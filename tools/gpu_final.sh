#!/bin/bash
# last check of a round: full GPU suite, smoke, the official bench line (+ GEMM table), one ncu --set full capture of the
# backward attention kernels and the item-table GEMM
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q --tb=short -p no:cacheprovider -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest -m gpu exit $? : $(tail -n 1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke.log)"
timeout 600 python bench.py --gemm-table gpurun_out/gemm_table.md > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-260 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err; head -20 gpurun_out/gemm_table.md
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph"
timeout 170 ncu --set full --clock-control none --import-source on -k regex:"attn_dq_pair|attn_dkv_pair" -c 2 -o gpurun_out/prof_attn_bwd $B > gpurun_out/ncu_b.log 2>&1
echo "== prof_attn_bwd exit $?"
ncu -i gpurun_out/prof_attn_bwd.ncu-rep --page raw --csv > gpurun_out/prof_attn_bwd.raw.csv 2>/dev/null
