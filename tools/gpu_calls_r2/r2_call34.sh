#!/usr/bin/env bash
# Round 2, GPU call 34 (1 GPU, the last minutes): smoke + the model tests on the committed tree (host-side change since call 33:
# HealpyGCNN(graph_builder=...))
set -u
mkdir -p gpurun_out
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1 > gpurun_out/r2ah_smoke.log
cat gpurun_out/r2ah_smoke.log
timeout 100 python -m pytest tests/test_gpu_model.py -q -m gpu 2>&1 | tail -1
