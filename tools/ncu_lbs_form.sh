#!/bin/bash
# One `ncu --set full` capture of the dominant kernel of a dense-LBS kernel form (B200_PROFILING.md recipe), plus the launch
# list of the same command.  Run on the GPU box:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/ncu_lbs_form.sh 3 4 lbs_fuseg_kernel r02a'
# args: skin form, blend form, kernel-name regex, tag for the output files (gpurun_out/<tag>_*)
set -u
SKIN=${1:-3}; BLEND=${2:-3}; KERNEL=${3:-lbs_fuseg_kernel}; TAG=${4:-form${SKIN}${BLEND}}
mkdir -p gpurun_out
export HB_LBS_SKIN=$SKIN HB_LBS_BLEND=$BLEND
# launch list (cold-cache, serialised: compare SHARES, not absolutes)
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python tools/run_lbs_once.py > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_summarize.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
# full capture of the second launch of the kernel (the first one is the warm-up)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s 1 -c 1 -o gpurun_out/${TAG}_set_full -f \
  python tools/run_lbs_once.py > gpurun_out/${TAG}_set_full.log 2>&1
ncu -i gpurun_out/${TAG}_set_full.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' > gpurun_out/${TAG}_set_full_key_metrics.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    names, units, vals = rows[0], rows[1], rows[-1]
    want = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__inst_executed_pipe_tensor', 'sm__pipe_tensor', 'l1tex__data_bank_conflicts_pipe_lsu', 'smsp__warp_issue_stalled', 'launch__registers_per_thread',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed')
    for n, u, v in zip(names, units, vals):
        if any(n.startswith(w) for w in want):
            print(f'{n:90s} {v:>18s} {u}')
PY
tail -n 12 gpurun_out/${TAG}_launches_summary.txt
head -n 40 gpurun_out/${TAG}_set_full_key_metrics.txt
