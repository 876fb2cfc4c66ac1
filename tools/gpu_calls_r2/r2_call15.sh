#!/usr/bin/env bash
# Round 2, GPU call 15 (8 GPUs): partitioned nside-1024 network at global batch 32: N = 1 (GPU 0 alone) and N = 8
set -u
mkdir -p gpurun_out
F="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-f-sweep"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --gpus 1 $F --no-configs > gpurun_out/r2o_part_n1.json 2> gpurun_out/r2o_part_n1.err
tail -c 300 gpurun_out/r2o_part_n1.err | grep -v sbi_flows
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 $F > gpurun_out/r2o_part_n8.json 2> gpurun_out/r2o_part_n8.err
python - <<'PY'
import json
for n in (1, 8):
    try:
        d = json.loads([l for l in open(f'gpurun_out/r2o_part_n{n}.json').read().strip().splitlines() if l.startswith('{')][-1])
        m = d.get('model_train_partitioned') or {}
        print(n, json.dumps({a: m.get(a) for a in ('value','ms_per_step','eager_ms_per_step','global_batch','peak_memory_GB','execution','time_split_ms','error','trace')})[:900])
        print('   graph', json.dumps(m.get('cuda_graph'))[:300])
        for c, v in (d.get('named_configs') or {}).items():
            print('  ', c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error')})[:300], json.dumps((v.get('cuda_graph') or {}).get('error')))
    except Exception as e:
        print(n, 'no json', e)
PY
