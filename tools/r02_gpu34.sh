#!/bin/bash
# Round 2, GPU call 34: what the driver runs at round end, on the final state: full GPU suite, smoke(), the default bench line (with
# its cpu_baseline, torch_cuda_port and roofline legs), the per-kernel profile of a step and the ncu launch list of a short bench run.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -12 | cut -c1-300) > gpurun_out/r03i_tests.txt
tail -3 gpurun_out/r03i_tests.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r03i_smoke.txt
tail -2 gpurun_out/r03i_smoke.txt
(timeout 900 python bench.py 2>gpurun_out/r03i_bench.err) > gpurun_out/r03i_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r03i_bench.json'))
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'vs_baseline', 'dtype')})
print('e2e', d['e2e']); print('roofline', {k: d['roofline'].get(k) for k in ('achieved', 'peak', 'frac', 'traffic', 'ms_per_launch', 'ms_dense_forward_call', 'timed')})
print('cpu_baseline', d.get('cpu_baseline')); print('port', d.get('torch_cuda_port'), d.get('vs_torch_cuda_port')); print('clocks', d.get('clocks'))
PY
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -34) > gpurun_out/r03i_profile_step.txt
head -12 gpurun_out/r03i_profile_step.txt | cut -c1-150
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r03i_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r03i_bench_under_ncu.log 2>&1
python tools/ncu_summarize.py gpurun_out/r03i_bench_launches.csv > gpurun_out/r03i_bench_launches_summary.txt 2>&1
head -14 gpurun_out/r03i_bench_launches_summary.txt | cut -c1-150
