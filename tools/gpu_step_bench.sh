#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gpu_step.py tests/test_gpu_finetune.py tests/test_gpu_vocab_parallel.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/$n.log 2>&1
  echo "== $n exit $? : $(tail -n 1 gpurun_out/$n.log)"; grep -E "^E  |FAILED|^\[" gpurun_out/$n.log | head -12
done
timeout 900 python bench.py --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline --gemm-table gpurun_out/gemm_table.md > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "== bench exit $?"; python - <<'P'
import json
try:
    j=json.load(open('gpurun_out/bench.json'))
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches','status_flags')}, 'e2e',j['e2e']['value'])
    print('gemm',j['roofline']['frac'],'share',j['roofline']['share_of_step'],'attn',j['roofline_attention']['fwd'],j['roofline_attention']['bwd'],j['roofline_attention']['share_of_step'],'step',j['step_roofline'])
except Exception as e: print('bench parse failed',e)
P
tail -n 5 gpurun_out/bench.err; grep -E "swiglu_bwd|f32\+= \| 5|f32 \| 5|bias_gelu|16384 \| 768 \| 8192" gpurun_out/gemm_table.md
