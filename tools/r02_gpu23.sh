#!/bin/bash
# Round 2, GPU call 23: the dense LBS pass (side stream, no consumer inside the closure) with MORE CTAs than SMs, so that it fills
# the SMs the main stream leaves idle instead of holding 148 SMs for 1 ms.
mkdir -p gpurun_out
for n in 0 296 592 1184 2368; do
  if [ $n = 0 ]; then unset HB_LBS_FUSEG_CTAS; else export HB_LBS_FUSEG_CTAS=$n; fi
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02w_ctas_$n.err) > gpurun_out/r02w_ctas_$n.json
  python - gpurun_out/r02w_ctas_$n.json $n <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('fuseg CTAs', sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'lbs alone ms', round(d['roofline']['ms_per_launch'], 3), d.get('result_check'))
PY
done
