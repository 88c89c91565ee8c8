#!/bin/bash
# Round 2, GPU call 4: reworked chain epilogue (st.async exchange, prefetched constants, parallel fences) + fused LBS default.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_closure.py tests/test_gpu_kernels.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/r02d_tests.txt
tail -4 gpurun_out/r02d_tests.txt
(timeout 200 python tools/chain_timeline.py 256 59 2>gpurun_out/r02d_chain_timeline.err) > gpurun_out/r02d_chain_timeline.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02d_chain_timeline.json'))
print('rollout fwd/bwd ms', d['rollout_fwd_ms'], d['rollout_bwd_ms'])
for k in ('fwd', 'bwd'):
    print(k, 'step period', d[k]['step_period_us_median'], {p: d[k][p].get('phase_total(0-11)', d[k][p].get('phase_total(0-3)')) for p in ('phase0', 'phase1', 'phase2', 'phase3', 'phase4')})
    print('   phase1', d[k]['phase1'])
    print('   glue', d[k]['phase4'], d[k]['handover_us'])
PY
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -42) > gpurun_out/r02d_profile_step.txt
cat gpurun_out/r02d_profile_step.txt | head -34
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02d_bench.err) > gpurun_out/r02d_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02d_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], d['step_breakdown_ms'])
PY
