#!/bin/bash
# Round 2, GPU call 10: whole GPU suite + bench after the try_wait suspend hint and the 2-rows-per-block GMM kernel.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -25 | cut -c1-400) > gpurun_out/r02j_gpu_tests.txt
tail -6 gpurun_out/r02j_gpu_tests.txt
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -36) > gpurun_out/r02j_profile_step.txt
head -14 gpurun_out/r02j_profile_step.txt | cut -c1-150
grep -E "gmm|fit_reduce|fit_losses" gpurun_out/r02j_profile_step.txt | cut -c1-120
(timeout 200 python tools/chain_timeline.py 256 59 2>gpurun_out/r02j_timeline.err) > gpurun_out/r02j_timeline.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02j_timeline.json'))
print('rollout fwd/bwd ms', round(d['rollout_fwd_ms'], 3), round(d['rollout_bwd_ms'], 3), 'periods', round(d['fwd']['step_period_us_median'], 2), round(d['bwd']['step_period_us_median'], 2))
print(' fwd ph1', d['fwd']['phase1']); print(' fwd glue', d['fwd']['phase4'], d['fwd']['handover_us']); print(' bwd glue', d['bwd']['phase4'], d['bwd']['handover_us'])
PY
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02j_bench.err) > gpurun_out/r02j_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02j_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['step_breakdown_ms'])
PY
