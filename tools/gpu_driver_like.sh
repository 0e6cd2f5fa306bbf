#!/bin/bash
# what the driver runs at round end, in its form: the GPU suite in ONE process, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/pytest_gpu_one_process.txt 2>&1
echo "== pytest -m gpu (one process) exit $? : $(grep -E 'passed|failed' gpurun_out/pytest_gpu_one_process.txt | tail -n 1) $(grep real gpurun_out/pytest_gpu_one_process.txt)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke.log)"
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
echo "== default bench exit $? $(grep real gpurun_out/bench_default.time)"; cut -c1-200 gpurun_out/bench_default.json
