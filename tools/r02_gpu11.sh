#!/bin/bash
# Round 2, GPU call 11: reduce-scatter as bulk DSMEM copies; two-level energy reduction.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_closure.py tests/test_gpu_kernels.py -q --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -12 | cut -c1-400) > gpurun_out/r02k_tests.txt
tail -5 gpurun_out/r02k_tests.txt
(timeout 200 python tools/chain_timeline.py 256 59 2>gpurun_out/r02k_timeline.err) > gpurun_out/r02k_timeline.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02k_timeline.json'))
print('rollout fwd/bwd ms', round(d['rollout_fwd_ms'], 3), round(d['rollout_bwd_ms'], 3), 'periods', round(d['fwd']['step_period_us_median'], 2), round(d['bwd']['step_period_us_median'], 2))
print(' fwd ph1', d['fwd']['phase1']); print(' bwd ph2', d['bwd']['phase2']); print(' bwd glue', d['bwd']['phase4'], d['bwd']['handover_us'])
PY
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02k_bench.err) > gpurun_out/r02k_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02k_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['step_breakdown_ms'], d['result_check'])
PY
(timeout 200 python tools/profile_step.py 256 60 2>&1 | grep -E "device busy|chain_kernel|fit_reduce|gmm|fuseg") | cut -c1-140
