#!/bin/bash
# Round 2, GPU call 26: the gradient-free dense LBS pass queued behind the launch of the reverse decoder chain (short CTAs on the SMs
# the chain leaves idle) vs right after the roll-out (HB_DENSE_EARLY=1); closure / run() tests first.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -15 | cut -c1-300) > gpurun_out/r03a_tests.txt
tail -3 gpurun_out/r03a_tests.txt
for mode in late early late; do
  if [ $mode = early ]; then export HB_DENSE_EARLY=1; else unset HB_DENSE_EARLY; fi
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r03a_bench_$mode.err) > gpurun_out/r03a_bench_$mode.json
  python - gpurun_out/r03a_bench_$mode.json $mode <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('dense pass', sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), d.get('step_breakdown_ms'), d.get('result_check'))
PY
done
unset HB_DENSE_EARLY
(timeout 200 python tools/chain_timeline.py 256 59 1965 2>gpurun_out/r03a_timeline.err) > gpurun_out/r03a_timeline.json
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -32) > gpurun_out/r03a_profile_step.txt
head -8 gpurun_out/r03a_profile_step.txt | cut -c1-150
