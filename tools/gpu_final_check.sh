#!/bin/bash
# One-shot hardware check of the alternative dense-LBS kernel forms + A/B bench + chamfer timing (bounded by timeouts).
mkdir -p gpurun_out
(timeout 80 python -m pytest tests/test_gpu_zz_lbs_forms.py -x -q -k "2-1-512 or rejects" 2>&1 | tail -6) > gpurun_out/t_forms_skin.log
(timeout 90 python -m pytest tests/test_gpu_zz_lbs_forms.py -x -q -k "not 2-1-512 and not rejects" 2>&1 | tail -12) > gpurun_out/t_forms_blend.log
(timeout 75 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin 2 2>gpurun_out/bench_skin2.err) > gpurun_out/bench_skin2.json
(timeout 75 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin 2 --lbs-blend 2 2>gpurun_out/bench_skin2_blend2.err) > gpurun_out/bench_skin2_blend2.json
(timeout 40 python tools/chamfer_time.py 240 2>&1 | tail -2) > gpurun_out/chamfer_time.json
tail -2 gpurun_out/t_forms_skin.log gpurun_out/t_forms_blend.log
for f in gpurun_out/bench_skin2.json gpurun_out/bench_skin2_blend2.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
cat gpurun_out/chamfer_time.json
