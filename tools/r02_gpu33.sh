#!/bin/bash
# Round 2, GPU call 33: + slot-less entries first in every group, their transform rows prefetched into L1.
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zz_lbs_forms.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/r03h_tests.txt
tail -4 gpurun_out/r03h_tests.txt
(timeout 200 python tools/lbs_forms_time.py --forms "3,5;3,1;1,1" --reps 10 2>gpurun_out/r03h_lbs_forms_time.err) > gpurun_out/r03h_lbs_forms_time.jsonl
cat gpurun_out/r03h_lbs_forms_time.jsonl | cut -c1-330
echo "HB_LBS_NO_SHAPE_ROWS=1"; HB_LBS_NO_SHAPE_ROWS=1 timeout 100 python tools/lbs_forms_time.py --forms "3,5" --reps 10 2>/dev/null | cut -c1-200 | tee gpurun_out/r03h_no_shape_rows.jsonl
bash tools/ncu_lbs_form.sh 3 5 lbs_fuseg_kernel r03h_fuseg35 > gpurun_out/r03h_ncu.log 2>&1
ncu -i gpurun_out/r03h_fuseg35_set_full.ncu-rep --page details > gpurun_out/r03h_fuseg35_set_full_details.txt 2>&1
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|L1/TEX Hit|Registers Per|Issued Instructions|dram__bytes|One or More Eligible" gpurun_out/r03h_fuseg35_set_full_details.txt gpurun_out/r03h_fuseg35_set_full_key_metrics.txt | head -20
(timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r03h_bench.err) > gpurun_out/r03h_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r03h_bench.json'))
print('bench ms/step', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), 'roofline', d['roofline']['frac'], d['roofline'].get('ms_per_launch'))
PY
