#!/bin/bash
# Round 2, GPU call 21: what the driver runs at round end - full GPU suite, smoke(), the default bench line (with its cpu_baseline,
# torch_cuda_port and roofline legs) and the reference arm with the driver's steps / warm-up.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -12 | cut -c1-300) > gpurun_out/r02u_tests.txt
tail -3 gpurun_out/r02u_tests.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r02u_smoke.txt
tail -2 gpurun_out/r02u_smoke.txt
(timeout 900 python bench.py 2>gpurun_out/r02u_bench.err) > gpurun_out/r02u_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02u_bench.json'))
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'vs_baseline', 'dtype')})
print('e2e', d['e2e']); print('roofline', {k: d['roofline'].get(k) for k in ('kernel', 'achieved', 'peak', 'frac', 'traffic')})
print('cpu_baseline', d.get('cpu_baseline')); print('port', d.get('torch_cuda_port'), d.get('vs_torch_cuda_port')); print('clocks', d.get('clocks'))
PY
(timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/r02u_ref.err) > gpurun_out/r02u_ref.json
cut -c1-700 gpurun_out/r02u_ref.json
