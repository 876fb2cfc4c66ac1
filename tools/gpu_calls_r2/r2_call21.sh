#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for c in C3_masked_survey C1_quick_start; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2u_launches_$c.csv python tools/profile_config.py $c > gpurun_out/r2u_prof_$c.log 2>&1
  tail -1 gpurun_out/r2u_prof_$c.log | cut -c1-200
done
