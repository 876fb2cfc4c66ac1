#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for c in C4_autoencoder C3_masked_survey; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2x_launches_$c.csv python tools/profile_config.py $c > gpurun_out/r2x_prof_$c.log 2>&1
  tail -1 gpurun_out/r2x_prof_$c.log | cut -c1-200
done
