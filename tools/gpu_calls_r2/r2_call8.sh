#!/usr/bin/env bash
# Round 2, GPU call 8: colsum rewrite + graph replays of the named configs; ncu evidence of the default build
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py tests/test_gpu_r2_networks.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2h_tests.log
tail -2 gpurun_out/r2h_tests.log
SECONDS=0; timeout 1500 python bench.py --no-e2e > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench wall seconds: $SECONDS"
grep -v "sbi_flows\|Warning\|warn" gpurun_out/r2h_bench.err | tail -c 400
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2h_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'kernels', json.dumps({k: round(v['ms'], 2) for k, v in d['kernels'].items()}))
    print('fsweep', json.dumps(d.get('fused_forward_narrow_layers')))
    for k in ('model_train','model_train_partitioned'):
        m = d.get(k) or {}
        print(k, json.dumps({a: m.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','error')})[:600])
    for c, v in (d.get('named_configs') or {}).items():
        print(c, json.dumps({a: v.get(a) for a in ('value','ms_per_step','eager_ms_per_step','execution','parity','error','trace')})[:700], json.dumps((v.get('cuda_graph') or {}).get('error')))
except Exception as e:
    print('no json', e)
PY
# ncu: launch list of the layer bench command (kernel shares) and one full capture of the dominant kernel (default build)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_layer_bench.csv python bench.py --steps 2 --warmup 1 --no-model --no-e2e --no-cpu-baseline --no-other-modes --no-f-sweep > gpurun_out/r2h_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lattice_conv2 -c 1 -f -o gpurun_out/r2h_prof_conv2_default python tools/profile_layer.py tf32 8 > gpurun_out/r2h_ncu_full.log 2>&1
tail -2 gpurun_out/r2h_ncu_full.log
