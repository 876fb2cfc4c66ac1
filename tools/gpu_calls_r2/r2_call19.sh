#!/usr/bin/env bash
# Round 2, GPU call 19 (2 GPUs): the whole GPU suite on the final tree (incl. ds_comm / partition tests), smoke
set -u
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -14
  echo "== smoke =="
  CUDA_VISIBLE_DEVICES=0 timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
} > gpurun_out/r2s_tests.log 2>&1
tail -10 gpurun_out/r2s_tests.log
