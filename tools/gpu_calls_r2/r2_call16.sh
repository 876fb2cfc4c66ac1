#!/usr/bin/env bash
# Round 2, GPU call 16: launch list of the HealpyGCNN training step (current default build)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2p_launches_model_train.csv python tools/profile_model.py > gpurun_out/r2p_model_prof.log 2>&1
tail -2 gpurun_out/r2p_model_prof.log
