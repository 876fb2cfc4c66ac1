#!/usr/bin/env bash
# Round 2, GPU call 29 (1 GPU): ncu --set full of the weight-gradient kernel at a NARROW shape (nside 128, batch 16, 16
# channels: 600 ns per 16-row stage whatever the stage carries - which role paces it?)
set -u
mkdir -p gpurun_out
NSIDE=128 FEATURES=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_tn_kernel -s 1 -c 1 \
  -f -o gpurun_out/r2ac_prof_umma_gemm_tn_narrow python tools/profile_layer.py tf32 16 bwd > gpurun_out/r2ac_prof.log 2>&1
tail -2 gpurun_out/r2ac_prof.log
ls -la gpurun_out/*.ncu-rep
