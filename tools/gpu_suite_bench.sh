#!/bin/bash
# full GPU suite (one process per file), smoke, the bench line with the per-shape GEMM table
mkdir -p gpurun_out
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/$n.log 2>&1
  echo "== $n exit $? : $(tail -n 1 gpurun_out/$n.log)"; grep -E "^E  |FAILED|^\[" gpurun_out/$n.log | head -12
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? : $(tail -n 1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps ${STEPS:-8} --warmup 3 --gemm-table gpurun_out/gemm_table.md > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; python - <<'P'
import json
try:
    j=json.load(open('gpurun_out/bench.json'))
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches','status_flags')}, 'e2e',j['e2e']['value'])
    print('gemm',j['roofline']['frac'],'share',j['roofline']['share_of_step'],'attn',j['roofline_attention']['fwd'],j['roofline_attention']['bwd'],j['roofline_attention']['share_of_step'],'step',j['step_roofline'])
    print('cpu',j.get('cpu_baseline'))
except Exception as e: print('bench parse failed',e)
P
tail -n 5 gpurun_out/bench.err; head -40 gpurun_out/gemm_table.md
