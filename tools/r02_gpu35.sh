#!/bin/bash
# Round 2, GPU call 35: GPU test of the shaped-template path of the fused LBS kernel; CTAs per SM of the dense pass behind the reverse chain.
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_zz_lbs_forms.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6) > gpurun_out/r03j_tests.txt
tail -3 gpurun_out/r03j_tests.txt
for c in 8 4 16 8; do
  (HB_DENSE_CTAS_PER_SM=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r03j_bench_ctas$c.err) > gpurun_out/r03j_bench_ctas$c.json
  python - gpurun_out/r03j_bench_ctas$c.json $c <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('dense CTAs per SM', sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3))
PY
done
