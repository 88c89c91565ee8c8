#!/bin/bash
# Round 2, GPU call 9: forward chain on fp16 planes inside the persistent kernel (precision tensor16), updated tests.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_closure.py tests/test_gpu_zz_run_e2e.py tests/test_gpu_zz_umma16.py -q --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -30 | cut -c1-500) > gpurun_out/r02i_tests.txt
tail -8 gpurun_out/r02i_tests.txt
for prec in tensor tensor16; do
  (timeout 200 python tools/chain_timeline.py 256 59 1965 $prec 2>gpurun_out/r02i_timeline_$prec.err) > gpurun_out/r02i_timeline_$prec.json
  python - gpurun_out/r02i_timeline_$prec.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(d['precision'], 'rollout fwd/bwd ms', round(d['rollout_fwd_ms'], 3), round(d['rollout_bwd_ms'], 3), 'fwd period', round(d['fwd']['step_period_us_median'], 2), 'bwd period', round(d['bwd']['step_period_us_median'], 2))
print('   fwd phase1', d['fwd']['phase1'])
PY
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision $prec 2>gpurun_out/r02i_bench_$prec.err) > gpurun_out/r02i_bench_$prec.json
  python - gpurun_out/r02i_bench_$prec.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('bench', d['impl_details']['precision'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['step_breakdown_ms'], d['result_check'])
PY
done
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -36) > gpurun_out/r02i_profile_step.txt
head -16 gpurun_out/r02i_profile_step.txt | cut -c1-150
