#!/bin/bash
# Round 2, GPU call 27: precision 'tensor16' (forward chain + forward prior on fp16 hi/lo planes) against 'tensor' on the final state.
mkdir -p gpurun_out
for prec in tensor tensor16 tensor tensor16; do
  (timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision $prec 2>>gpurun_out/r03b_bench_$prec.err) > gpurun_out/r03b_bench_$prec.json
  python - gpurun_out/r03b_bench_$prec.json $prec <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), d.get('result_check'))
PY
done
