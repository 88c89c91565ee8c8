#!/bin/bash
# Round 2, GPU call 36 (2 GPUs): N=1 and N=2 bench lines of the final state back to back.
# exchange inside the CUDA graph with the round's new kernels).
mkdir -p gpurun_out
(timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r03k_n1.err) > gpurun_out/r03k_n1.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r03k_n2.err) > gpurun_out/r03k_n2.json
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads([l for l in open(f'gpurun_out/r03k_n{n}.json') if l.startswith('{')][-1])
        print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'], 'per_step', d.get('per_step_ms', {}).get('by_rank_median'))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/r03k_n{n}.err').read()[-1500:])
PY
