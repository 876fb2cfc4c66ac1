#!/usr/bin/env bash
# Round 2, GPU call 32 (2 GPUs): the two-GPU tests on the final tree (per-pixel lattice validity, patch kernel, new
# weight-gradient producers under the partitioned layers), then the 2-rank bench line
set -u
mkdir -p gpurun_out
S=$SECONDS
timeout 600 python -m pytest tests/test_gpu_partition.py -q -m gpu 2>&1 | tail -6 > gpurun_out/r2af_tests_partition_2gpu.log
tail -2 gpurun_out/r2af_tests_partition_2gpu.log
echo "tests: $((SECONDS-S)) s"; S=$SECONDS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-configs --no-f-sweep --no-other-modes > gpurun_out/r2af_bench_2gpu.json 2> gpurun_out/r2af_bench_2gpu.err
echo "bench: $((SECONDS-S)) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2af_bench_2gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'))
mt = d.get('model_train') or {}
print('model_train', mt.get('value'), mt.get('ms_per_step'))
mp = d.get('model_train_partitioned') or {}
print('partitioned', json.dumps({k: mp.get(k) for k in ('value', 'ms_per_step', 'execution')})[:300])
print('partition_parity', json.dumps(d.get('partition_parity') or mp.get('partition_parity'))[:400])
PY
