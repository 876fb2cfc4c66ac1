#!/usr/bin/env bash
# Round 2, GPU call 14 (8 GPUs): the whole default bench under torchrun, as the driver's scaling run launches it
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench_8gpu.err
echo "bench wall seconds: $SECONDS rc=$?"
grep -v "sbi_flows\|Warning\|warn\|OMP_NUM\|\*\*\*" gpurun_out/r2n_bench_8gpu.err | tail -c 1200
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2n_bench_8gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', json.dumps(d.get('e2e'))[:400])
    for k in ('model_train','model_train_partitioned'):
        m = d.get(k) or {}
        print(k, json.dumps({a: m.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','time_split_ms','limiting_collective','host_prep_s','error','trace')})[:900])
        print('   graph', json.dumps(m.get('cuda_graph'))[:300])
    print('parity', json.dumps(d.get('partition_parity'))[:400])
    for c, v in (d.get('named_configs') or {}).items():
        print(c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error','trace')})[:500])
except Exception as e:
    print('no json', e)
PY
