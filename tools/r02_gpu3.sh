#!/bin/bash
# Round 2, third GPU call: per-phase timeline of the persistent chain, fixed chain tests + xbatch golden, bench of the other
# BASELINE configs (C4: RGB B=16/GPU, C5: AMASS key-vertex sweep), reference arm with the driver's --steps/--warmup.
mkdir -p gpurun_out
(timeout 200 python tools/chain_timeline.py 256 59 2>gpurun_out/r02c_chain_timeline.err) > gpurun_out/r02c_chain_timeline.json
cat gpurun_out/r02c_chain_timeline.json
(timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_closure.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/r02c_chain_closure_tests.txt
tail -6 gpurun_out/r02c_chain_closure_tests.txt
for b in 32 64 128 256 512; do
  (timeout 200 python bench.py --config amass --batch $b --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin 3 --lbs-blend 5 2>gpurun_out/r02c_bench_amass_b$b.err) > gpurun_out/r02c_bench_amass_b$b.json
done
(timeout 200 python bench.py --config rgb --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin 3 --lbs-blend 5 2>gpurun_out/r02c_bench_rgb_b16.err) > gpurun_out/r02c_bench_rgb_b16.json
(timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/r02c_ref_arm.err) > gpurun_out/r02c_ref_arm.json
(timeout 400 python bench.py --steps 20 --warmup 5 --lbs-skin 3 --lbs-blend 5 2>gpurun_out/r02c_bench_full.err) > gpurun_out/r02c_bench_full.json
for f in gpurun_out/r02c_bench*.json gpurun_out/r02c_ref_arm.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), (d.get('cpu_baseline') or {}).get('sample'), d.get('torch_cuda_port'), d.get('vs_torch_cuda_port'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
