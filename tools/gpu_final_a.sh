#!/bin/bash
# round-end evidence, part A: full GPU suite (one process per file), smoke, the bench line, the reference arm, per-kernel roofline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used,power.limit --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/pytest_gpu.txt
for f in tests/test_gpu_*.py tests/test_torch_ext.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/$n.log 2>&1
  echo "== $n exit $? : $(tail -n 1 gpurun_out/$n.log)" | tee -a gpurun_out/pytest_gpu.txt; grep -E "^E  |FAILED" gpurun_out/$n.log | head -8
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke.log)" | tee -a gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 --gemm-table gpurun_out/gemm_table.md > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; cut -c1-400 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "== reference exit $?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 600 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; tail -n 36 gpurun_out/kernel_roofline.log
