#!/bin/bash
# Round 2, GPU call 19: GMM factor staging with batched loads, d-feature pass of the sparse skinning reverse on the packed columns: full GPU suite, bench A/B (one tile per CTA vs persistent),
# per-kernel profile of the step, ncu --set full of the persistent GEMM.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -25 | cut -c1-400) > gpurun_out/r02s_tests.txt
tail -4 gpurun_out/r02s_tests.txt
for mode in persistent; do
  if [ $mode = one_tile ]; then export HB_GEMM_ONE_TILE=1; else unset HB_GEMM_ONE_TILE; fi
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02s_bench_$mode.err) > gpurun_out/r02s_bench_$mode.json
  python - gpurun_out/r02s_bench_$mode.json $mode <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('bench', sys.argv[2], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('step_breakdown_ms'), 'launches/step', d.get('gpu_launches_per_step'), d.get('result_check'))
PY
done
unset HB_GEMM_ONE_TILE
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -40) > gpurun_out/r02s_profile_step.txt
head -24 gpurun_out/r02s_profile_step.txt | cut -c1-150
