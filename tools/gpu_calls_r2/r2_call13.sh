#!/usr/bin/env bash
# Round 2, GPU call 13: model steps (narrow layers: a drain every 2 items) with the drain on the IO warps
set -u
mkdir -p gpurun_out
for name in default iodrain iodrainsymw symw; do
  V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ "$name" = default ] && V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200.so
  DEEPSPHERE_LIB=$V timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs 2>/dev/null | grep '^{' > gpurun_out/r2m_bench_$name.json
done
python - <<'PY'
import json
for name in ('default', 'iodrain', 'iodrainsymw', 'symw'):
    try:
        d = json.load(open(f'gpurun_out/r2m_bench_{name}.json'))
        print(name, 'layer', round(d['ms_per_step'], 2), 'fwd', round(d['kernels']['forward']['ms'], 2), 'F16/32', [round(v['ms'], 2) for v in d['fused_forward_narrow_layers'].values()],
              'model_train', round(d['model_train']['ms_per_step'], 3), 'partitioned', round(d['model_train_partitioned']['ms_per_step'], 2))
    except Exception as e:
        print(name, 'failed', e)
PY
