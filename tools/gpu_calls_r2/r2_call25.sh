#!/usr/bin/env bash
# Round 2, GPU call 25 (2 GPUs): the whole GPU suite on the final tree (incl. the two-GPU tests), smoke, 2-rank bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2y_tests_2gpu.log
tail -3 gpurun_out/r2y_tests_2gpu.log
echo "== smoke ==" >> gpurun_out/r2y_tests_2gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 >> gpurun_out/r2y_tests_2gpu.log
tail -1 gpurun_out/r2y_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2y_bench_2gpu.json 2> gpurun_out/r2y_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2y_bench_2gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
mt = d.get('model_train') or {}
print('model_train', mt.get('value'), mt.get('ms_per_step'))
mp = d.get('model_train_partitioned') or {}
print('partitioned', json.dumps({k: mp.get(k) for k in ('value', 'ms_per_step', 'execution', 'batch', 'parity')})[:400])
PY
