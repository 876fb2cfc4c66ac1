#!/usr/bin/env bash
# Round 2, GPU call 6: diagnose the CUDA-graph capture of the training step; chained K=10 lattice passes; network parity
set -u
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python bench.py --model-only --model-graph > gpurun_out/r2f_graph.out 2> gpurun_out/r2f_graph.err
echo "graph rc=$?"; tail -c 1500 gpurun_out/r2f_graph.err; tail -c 300 gpurun_out/r2f_graph.out
timeout 600 python -m pytest tests/test_gpu_r2_networks.py "tests/test_gpu_lattice.py::test_lattice_forward_backward_matches_oracle" -q -m gpu 2>&1 | tail -15 > gpurun_out/r2f_tests.log
tail -8 gpurun_out/r2f_tests.log
