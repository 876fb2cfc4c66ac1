#!/usr/bin/env bash
# Round 2, GPU call 9 (2 GPUs): partitioned nside-1024 model at global batch 16: N = 1 vs N = 2, graph replays with NCCL inside
set -u
mkdir -p gpurun_out
F="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs --no-f-sweep --part-batch 16"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --gpus 1 $F > gpurun_out/r2i_part_n1.json 2> gpurun_out/r2i_part_n1.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $F > gpurun_out/r2i_part_n2.json 2> gpurun_out/r2i_part_n2.err
grep -v "sbi_flows\|Warning\|warn" gpurun_out/r2i_part_n2.err | tail -c 500
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads([l for l in open(f'gpurun_out/r2i_part_n{n}.json').read().strip().splitlines() if l.startswith('{')][-1])
        m = d.get('model_train_partitioned') or {}
        print(n, json.dumps({a: m.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','time_split_ms','error','trace')})[:900])
        print('   graph', json.dumps(m.get('cuda_graph'))[:400])
        print('   parity', json.dumps(d.get('partition_parity'))[:400])
        print('   model_train', json.dumps({a: (d.get('model_train') or {}).get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution')}))
    except Exception as e:
        print(n, 'no json', e)
PY
