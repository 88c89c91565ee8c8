#!/bin/bash
# Round 2, GPU call 30: what paces the fused LBS kernel - measurement switches (results are wrong with them on; timing only).
mkdir -p gpurun_out
for d in 0 1 2 3 4; do
  echo "HB_LBS_FUSEG_DBG=$d" >> gpurun_out/r03e_dbg.txt
  HB_LBS_FUSEG_DBG=$d timeout 100 python tools/lbs_forms_time.py --forms "3,5" --reps 10 2>/dev/null | cut -c1-160 >> gpurun_out/r03e_dbg.txt
done
cat gpurun_out/r03e_dbg.txt
