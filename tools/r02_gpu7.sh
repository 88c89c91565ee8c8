#!/bin/bash
# Round 2, GPU call 7: 16-warp epilogue of the fused dense LBS kernel.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_zz_lbs_forms.py tests/test_gpu_kernels.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/r02f_tests.txt
tail -4 gpurun_out/r02f_tests.txt
(timeout 120 python tools/lbs_forms_time.py --forms "3,5;3,1;3,4" --peak-gbs 6490.5 2>gpurun_out/r02f_lbs_forms_time.err) > gpurun_out/r02f_lbs_forms_time.jsonl
cat gpurun_out/r02f_lbs_forms_time.jsonl
(HB_LBS_FUSEG_DIRECT=1 timeout 120 python tools/lbs_forms_time.py --forms "3,5" --peak-gbs 6490.5 2>/dev/null) > gpurun_out/r02f_lbs_forms_time_direct.jsonl
cat gpurun_out/r02f_lbs_forms_time_direct.jsonl
bash tools/ncu_lbs_form.sh 3 5 lbs_fuseg_kernel r02f_fuseg35 > gpurun_out/r02f_fuseg35.log 2>&1
ncu -i gpurun_out/r02f_fuseg35_set_full.ncu-rep --page details 2>/dev/null | grep -E "Duration|Issue Slots Busy|Registers Per|Eligible Warps|Executed Instructions |Warp Cycles Per Issued|DRAM Throughput|Achieved Occupancy|Executed Ipc"
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02f_bench.err) > gpurun_out/r02f_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02f_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], d['step_breakdown_ms'])
PY
