#!/bin/bash
# Round 2, GPU call 37, final state of the round: weak scaling at 8 GPUs (neighbour halo) and the N=1 line of the same box, driver-style step counts.
mkdir -p gpurun_out
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r03l_n8.err) > gpurun_out/r03l_n8.json
(timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r03l_n1.err) > gpurun_out/r03l_n1.json
python - <<'PY'
import json
v = {}
for n in (1, 8):
    try:
        d = json.loads([l for l in open(f'gpurun_out/r03l_n{n}.json') if l.startswith('{')][-1])
        v[n] = d['value']
        print(n, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), d['per_step_ms'], d['clocks'])
    except Exception as e:
        print(n, 'unreadable', e); print(open(f'gpurun_out/r03l_n{n}.err').read()[-1200:])
if len(v) == 2:
    print('weak-scaling efficiency', v[8] / (8 * v[1]))
PY
