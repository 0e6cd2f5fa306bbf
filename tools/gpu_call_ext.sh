#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_torch_ext.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/test_torch_ext.log 2>&1
echo "== torch_ext exit $? : $(tail -n 1 gpurun_out/test_torch_ext.log)"; grep -E "^E  |FAILED" gpurun_out/test_torch_ext.log | head -12
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "swiglu or gemm" -q --tb=short -p no:cacheprovider -m gpu 2>&1 | tail -2
for c in C2 C3 C5; do
  timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  echo "== bench $c exit $?"; python - <<P
import json
try:
    j=json.load(open('gpurun_out/bench_$c.json'))
    print('$c', round(j['ms_per_step'],2),'ms', round(j['value']), 'e2e', round(j['e2e']['value']), 'gemm', round(j['roofline']['frac'],3), 'attn fwd/bwd', round(j['roofline_attention']['fwd']['frac'],3), round(j['roofline_attention']['bwd']['frac'],3), 'us', round(j['roofline_attention']['fwd']['us_per_launch']), round(j['roofline_attention']['bwd']['us_per_launch']), 'flags', j['status_flags'])
except Exception as e: print('parse failed', e)
P
  tail -n 3 gpurun_out/bench_$c.err
done
