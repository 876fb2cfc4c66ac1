#!/usr/bin/env bash
# Round 2, GPU call 5 (2 GPUs): the whole bench under torchrun - partition parity (N ranks == 1 rank) and the nside-1024
# sphere-partitioned HealpyGCNN - plus the 2-GPU NCCL parity tests that a 1-GPU box skips
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_partition.py -q -m gpu 2>&1 | tail -6 > gpurun_out/r2e_tests_2gpu.log
tail -3 gpurun_out/r2e_tests_2gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
tail -c 800 gpurun_out/r2e_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2e_bench_2gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
    for k in ('value','ms_per_step','e2e','model_train_partitioned','partition_parity'):
        print(k, json.dumps(d.get(k))[:1200])
    print('model_train', json.dumps({k:v for k,v in (d.get('model_train') or {}).items() if k in ('value','ms_per_step','error')}))
except Exception as e:
    print('no json', e)
PY
