#!/bin/bash
# Round 2: weak scaling at 8 GPUs with per-step stamps - neighbour halo (default) vs the round-1 all_gather/all_reduce vs no coupling.
mkdir -p gpurun_out
run() { # name, env..., extra flags
  name=$1; shift
  (env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline $EXTRA 2>gpurun_out/r02h_$name.err) > gpurun_out/r02h_$name.json
  python - "gpurun_out/r02h_$name.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), d['per_step_ms'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
EXTRA="" run neighbour HB_HALO=neighbour
EXTRA="" run allgather HB_HALO=allgather
EXTRA="--no-halo" run nohalo HB_HALO=neighbour
(timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02h_n1.err) > gpurun_out/r02h_n1.json
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02h_n1.json') if l.startswith('{')][-1])
print('N=1 value', round(d['value']), 'ms/step', d['ms_per_step'], d['per_step_ms'])
PY
