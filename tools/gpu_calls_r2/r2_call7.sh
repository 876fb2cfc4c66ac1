#!/usr/bin/env bash
# Round 2, GPU call 7: the whole default bench at N = 1 (named configs, graph replays, partitioned model)
set -u
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
echo "bench wall seconds: $SECONDS"
grep -v "sbi_flows\|Warning\|warn" gpurun_out/r2g_bench_default.err | tail -c 600
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2g_bench_default.json').read().strip().splitlines() if l.startswith('{')][-1])
    for k in ('value','ms_per_step','parity'):
        print(k, json.dumps(d.get(k))[:400])
    for k in ('model_train','model_train_partitioned'):
        m = d.get(k) or {}
        print(k, json.dumps({a: m.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error','trace','time_split_ms')})[:900])
        print('   graph', json.dumps(m.get('cuda_graph'))[:500])
    for c, v in (d.get('named_configs') or {}).items():
        print(c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','roofline','cpu_baseline','parity','error','trace')})[:1000])
except Exception as e:
    print('no json', e)
PY
