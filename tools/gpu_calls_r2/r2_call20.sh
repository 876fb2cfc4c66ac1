#!/usr/bin/env bash
# Round 2, GPU call 20: launch list of the partitioned nside-1024 step at N = 1 (time split evidence); default bench JSON
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2t_launches_partitioned_n1.csv python tools/profile_partitioned.py > gpurun_out/r2t_part_prof.log 2>&1
tail -2 gpurun_out/r2t_part_prof.log | cut -c1-400
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err
echo "bench wall seconds: $SECONDS; stdout lines: $(wc -l < gpurun_out/r2t_bench_default.json)"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2t_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'parity ok', d['parity']['ok'], 'roofline', d['roofline']['frac'])
m = d['model_train_partitioned']; print('partitioned', m['ms_per_step'], m['global_batch'], m.get('halo_bytes_per_step_per_rank'), m['execution'])
print('model_train', d['model_train']['ms_per_step'], d['model_train']['value'])
PY
