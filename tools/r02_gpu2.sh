#!/bin/bash
# Round 2, second GPU call: first hardware run of the persistent decoder-chain kernel.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_chain.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/r02b_chain_tests.txt
tail -12 gpurun_out/r02b_chain_tests.txt
(timeout 300 python tools/chain_accuracy.py 2>gpurun_out/r02b_chain_accuracy.err) > gpurun_out/r02b_chain_accuracy.jsonl
cat gpurun_out/r02b_chain_accuracy.jsonl
if true; then
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02b_bench.err) > gpurun_out/r02b_bench.json
  (HB_CHAIN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02b_bench_legacy.err) > gpurun_out/r02b_bench_legacy.json
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --lbs-skin 3 --lbs-blend 5 2>gpurun_out/r02b_bench_s3b5.err) > gpurun_out/r02b_bench_s3b5.json
  (timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/r02b_gpu_tests.txt
  # launch list of one step + a --set full capture of the chain kernel (forward and reverse launch)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 2000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02b_launches.log 2>&1
  python tools/ncu_summarize.py gpurun_out/r02b_launches.csv > gpurun_out/r02b_launches_summary.txt 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 2 -o gpurun_out/r02b_chain_kernel_set_full -f python tools/run_rollout_once.py 59 > gpurun_out/r02b_chain_set_full.log 2>&1
  (timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -40) > gpurun_out/r02b_profile_step.txt
fi
for f in gpurun_out/r02b_bench*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'LBS ms', round(d['roofline']['ms_per_launch'], 3), d.get('step_breakdown_ms'), d.get('result_check'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
tail -5 gpurun_out/r02b_gpu_tests.txt
head -20 gpurun_out/r02b_launches_summary.txt
