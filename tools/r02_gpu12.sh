#!/bin/bash
# Round 2, GPU call 12: glue prefetch of z / prior d-xin rows; MotionOptimizer.run end to end.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_closure.py -q --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -6 | cut -c1-300) > gpurun_out/r02l_tests.txt
tail -3 gpurun_out/r02l_tests.txt
(timeout 200 python tools/chain_timeline.py 256 59 2>gpurun_out/r02l_timeline.err) > gpurun_out/r02l_timeline.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02l_timeline.json'))
print('rollout fwd/bwd ms', round(d['rollout_fwd_ms'], 3), round(d['rollout_bwd_ms'], 3), 'periods', round(d['fwd']['step_period_us_median'], 2), round(d['bwd']['step_period_us_median'], 2))
print(' fwd glue', d['fwd']['phase4'], d['fwd']['handover_us']); print(' bwd glue', d['bwd']['phase4'], d['bwd']['handover_us'])
PY
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02l_bench.err) > gpurun_out/r02l_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02l_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['step_breakdown_ms'])
PY
(timeout 600 python tools/time_run.py 64 2>gpurun_out/r02l_time_run_b64.err) > gpurun_out/r02l_time_run_b64.jsonl
cat gpurun_out/r02l_time_run_b64.jsonl
(timeout 900 python tools/time_run.py 256 2>gpurun_out/r02l_time_run_b256.err) > gpurun_out/r02l_time_run_b256.jsonl
cat gpurun_out/r02l_time_run_b256.jsonl
tail -3 gpurun_out/r02l_time_run_b256.err
