#!/bin/bash
# First GPU call of the next round (≈22 min of box time): put the kernel forms that round 1 verified only on the CPU
# emulation onto the B200, time them, and decide the defaults from the STEP time.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_final_check.sh'
mkdir -p gpurun_out
# 1. correctness of the opt-in forms (asserts humor_lbs_forms_used == requested and a last-bit difference vs form 1)
(HB_TEST_UNVERIFIED=1 timeout 180 python -m pytest tests/test_gpu_zz_lbs_forms.py -x -q 2>&1 | tail -12) > gpurun_out/t_forms.log
# 2. stand-alone dense LBS forward per form (ms, GB/s, forms used, bitwise difference vs form 1)
(timeout 120 python tools/lbs_forms_time.py --peak-gbs 6490.5 2>gpurun_out/lbs_forms_time.err) > gpurun_out/lbs_forms_time.jsonl
# 3. step time with each candidate (the dense pass runs on a side stream under the decoder chain: persistent CTAs may delay it)
for cfg in "2 1" "2 2" "2 3" "3 1" "3 3" "3 4" "3 5"; do set -- $cfg
  (timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin $1 --lbs-blend $2 2>gpurun_out/bench_s$1b$2.err) > gpurun_out/bench_s$1b$2.json
done
# 3b. skin form 3 holds every SM it runs on (223 KB of shared memory, all of TMEM) for the whole pass: with fewer persistent
#     CTAs the decoder chain of the main stream keeps some SMs
(HB_LBS_FUSEG_CTAS=100 timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin 3 --lbs-blend 3 2>gpurun_out/bench_s3b3_100.err) > gpurun_out/bench_s3b3_100.json
# 3c. decoder chain: weight tiles requested before the programmatic-dependent-launch wait (default forms otherwise)
(HB_UMMA_PREFETCH_B=1 timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_prefetchb.err) > gpurun_out/bench_prefetchb.json
(timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_default.err) > gpurun_out/bench_default.json
(HB_UMMA_PREFETCH_B=1 timeout 200 python -m pytest tests/test_gpu_umma.py tests/test_gpu_kernels.py tests/test_gpu_closure.py -x -q 2>&1 | tail -4) > gpurun_out/t_prefetchb.log
# 3d. the fp16 hi/lo GEMM (4-byte operand elements): parity, then accuracy + in-situ kernel time against the 3xTF32 GEMM
(HB_TEST_UNVERIFIED=1 timeout 180 python -m pytest tests/test_gpu_zz_umma16.py -x -q 2>&1 | tail -4) > gpurun_out/t_umma16.log
(timeout 120 python tools/umma16_probe.py 2>gpurun_out/umma16_probe.err) > gpurun_out/umma16_probe.jsonl
# 3e. precision 'tensor16': forward decoder chain on 4-byte operand elements - parity against the reference fixtures, then the step time
(timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision tensor16 2>gpurun_out/bench_tensor16.err) > gpurun_out/bench_tensor16.json
# 4. the tests that round 1 could only run on the emulation
(HB_TEST_UNVERIFIED=1 timeout 150 python -m pytest tests/test_gpu_zz_stage12.py tests/test_gpu_zz_run_e2e.py -x -q 2>&1 | tail -6) > gpurun_out/t_stage12_e2e.log
tail -n 3 gpurun_out/t_forms.log gpurun_out/t_stage12_e2e.log gpurun_out/t_prefetchb.log gpurun_out/t_umma16.log
cat gpurun_out/lbs_forms_time.jsonl
cat gpurun_out/umma16_probe.jsonl
for f in gpurun_out/bench_s*.json gpurun_out/bench_default.json gpurun_out/bench_prefetchb.json gpurun_out/bench_tensor16.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), 'LBS ms', round(d['roofline']['ms_per_launch'], 3), 'frac', round(d['roofline']['frac'], 4))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
