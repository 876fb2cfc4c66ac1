#!/usr/bin/env bash
# Builds the A/B libraries of the fused-kernel experiments (run HERE before a gpurun call: the .so files travel with the
# snapshot).  Each is the same source with preprocessor switches; load one with DEEPSPHERE_LIB=<path>.
set -e
B=deepsphere-cosmo-tf2_b200/build.py
python $B                                                          # default
python $B --variant fence   -DC2_FENCE_BY_ISSUER=1                 # proxy fence by the UMMA-issuing lane
python $B --variant r216    -DC2_REGS_COMPUTE=216 -DC2_REGS_IO=40   # default blocks, 8 more registers for the compute warps
python $B --variant br2     -DC2_BR=2                              # 2x3 pixel blocks, 6 compute warps, 136 / 48 registers
python $B --variant br2r144 -DC2_BR=2 -DC2_REGS_COMPUTE=144 -DC2_REGS_IO=40
python $B --variant split   -DC2_SPLIT_BAR=1                       # hop_ready / hop_full: the proxy fence leaves the hop chain
python $B --variant noprobe -DC2_PROBE=0                            # timeline probe compiled out (-5 % instructions)
python $B --variant epipipe -DC2_EPI_PIPE=1                         # tensor-memory drain with the next load in flight
python $B --variant symw    -DC2_SYMW=1                            # symmetric in-block weights: 61 instead of 81 registers
python $B --variant br2symw -DC2_BR=2 -DC2_SYMW=1 -DC2_REGS_COMPUTE=144 -DC2_REGS_IO=40
python $B --variant br2symwepi -DC2_BR=2 -DC2_SYMW=1 -DC2_EPI_PIPE=1 -DC2_REGS_COMPUTE=144 -DC2_REGS_IO=40
python $B --variant br3all -DC2_SYMW=1 -DC2_EPI_PIPE=1 -DC2_PROBE=0 -DC2_SPLIT_BAR=1   # everything that keeps the measured block shape
python $B --variant br2all -DC2_BR=2 -DC2_SYMW=1 -DC2_EPI_PIPE=1 -DC2_PROBE=0 -DC2_SPLIT_BAR=1 -DC2_REGS_COMPUTE=144 -DC2_REGS_IO=40
python $B --variant br2allfence -DC2_BR=2 -DC2_SYMW=1 -DC2_EPI_PIPE=1 -DC2_PROBE=0 -DC2_SPLIT_BAR=1 -DC2_REGS_COMPUTE=144 -DC2_REGS_IO=40 -DC2_FENCE_BY_ISSUER=1
ls -la deepsphere-cosmo-tf2_b200/lib/*.so
