#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/part_layer_times.py 8 2>/dev/null | grep '^{' > gpurun_out/r2j_layers_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/part_layer_times.py 8 2>/dev/null | grep '^{' > gpurun_out/r2j_layers_n2.json
python - <<'PY'
import json
a = json.load(open('gpurun_out/r2j_layers_n1.json')); b = json.load(open('gpurun_out/r2j_layers_n2.json'))
print('fwd', a['fwd_total'], b['fwd_total'], 'bwd', a['bwd_total'], b['bwd_total'])
for x, y in zip(a['layers'], b['layers']):
    print(f"{x['layer']:28s} fwd {x['fwd_ms']:8.3f} -> {y['fwd_ms']:8.3f}   bwd {x['bwd_ms']:8.3f} -> {y['bwd_ms']:8.3f}")
PY
