#!/usr/bin/env bash
# Round 2, GPU call 12: accumulator drain on the IO warps; the partitioned step with the streaming pseudo-conv kernels
set -u
mkdir -p gpurun_out
for name in iodrain iodrainsymw; do
  V=$PWD/deepsphere-cosmo-tf2_b200/lib/libdeepsphere_b200_$name.so
  [ -f "$V" ] || continue
  {
    DEEPSPHERE_LIB=$V timeout 300 python -m pytest tests/test_gpu_lattice.py -q -m gpu -k "fused or conv2" 2>&1 | tail -12
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 2>&1 | grep RESULT
    DEEPSPHERE_LIB=$V timeout 200 python tools/bench_fwd.py tf32 32 bwd 2>&1 | grep RESULT
  } > gpurun_out/r2l_variant_$name.log 2>&1
  echo "$name: $(grep -E 'passed|failed' gpurun_out/r2l_variant_$name.log | tail -1) | $(grep RESULT gpurun_out/r2l_variant_$name.log | cut -c1-100 | tr '\n' '|')"
  grep -m2 -E "Error|error" gpurun_out/r2l_variant_$name.log | cut -c1-300
done
for sk in 0 1; do
  DEEPSPHERE_SKINNY=$sk timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs --no-f-sweep 2>/dev/null | grep '^{' > gpurun_out/r2l_bench_skinny$sk.json
done
python - <<'PY'
import json
for sk in (0, 1):
    d = json.load(open(f'gpurun_out/r2l_bench_skinny{sk}.json'))
    for k in ('model_train', 'model_train_partitioned'):
        m = d[k]
        print('skinny', sk, k, m.get('ms_per_step'), m.get('eager_ms_per_step'), m.get('final_loss'), (m.get('cuda_graph') or {}).get('validated'))
PY
