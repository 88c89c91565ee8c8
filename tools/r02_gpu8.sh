#!/bin/bash
# Round 2, GPU call 8: full-group 16-warp LBS epilogue; new GPU tests (native L-BFGS, chain oracle errors); tensor-mode tolerances.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_zz_lbs_forms.py tests/test_gpu_kernels.py tests/test_gpu_lbfgs.py tests/test_gpu_chain.py -q --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -40 | cut -c1-400) > gpurun_out/r02g_tests.txt
tail -12 gpurun_out/r02g_tests.txt
(timeout 120 python tools/lbs_forms_time.py --forms "3,5;3,4" --peak-gbs 6490.5 2>gpurun_out/r02g_lbs_forms_time.err) > gpurun_out/r02g_lbs_forms_time.jsonl
cat gpurun_out/r02g_lbs_forms_time.jsonl
bash tools/ncu_lbs_form.sh 3 5 lbs_fuseg_kernel r02g_fuseg35 > gpurun_out/r02g_fuseg35.log 2>&1
ncu -i gpurun_out/r02g_fuseg35_set_full.ncu-rep --page details 2>/dev/null | grep -E "Duration|Issue Slots Busy|Registers Per|Eligible Warps|Executed Instructions |Warp Cycles Per Issued|DRAM Throughput|Achieved Occupancy|Executed Ipc"
(timeout 600 python tools/tensor_mode_tolerances.py 2>gpurun_out/r02g_tolerances.err) > gpurun_out/r02g_tolerances.jsonl
cat gpurun_out/r02g_tolerances.jsonl
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02g_bench.err) > gpurun_out/r02g_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02g_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], d['step_breakdown_ms'], d['per_step_ms'])
PY
