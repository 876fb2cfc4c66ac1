#!/usr/bin/env bash
# Builds A/B libraries of the fused-kernel experiments (run HERE before a gpurun call: the .so files travel with the
# snapshot).  Each is the same source with preprocessor switches; load one with DEEPSPHERE_LIB=<path>.  The round-2
# measurements of all of them: profiles/r2_fused_kernel_variants.txt, DESIGN.md section 3.
set -e
B=deepsphere-cosmo-tf2_b200/build.py
python $B                                                            # default (timeline probe compiled out)
python $B --variant probe   -DC2_PROBE=1                              # clock64 hop timeline (DEEPSPHERE_CONV2_DEBUG=1)
python $B --variant split   -DC2_SPLIT_BAR=1                          # hop_ready / hop_full: proxy fence off the hop chain
python $B --variant symw    -DC2_SYMW=1                               # symmetric in-block weights: 61 instead of 81 registers
python $B --variant loop    -DC2_LOOP=1                               # hops as a loop over one body per buffer parity
python $B --variant bc6     -DC2_CPT=2 -DC2_BC=6 -DC2_SYMW=1          # 3 x 6 pixels x 2 channels per thread, scalar diagonal
python $B --variant bc6loop -DC2_CPT=2 -DC2_BC=6 -DC2_SYMW=1 -DC2_LOOP=1
python $B --variant cpt2    -DC2_CPT=2 -DC2_SYMW=1 -DC2_CDIAG=1        # 3 x 3 x 2 channels, 8 compute warps per CTA (spills)
python $B --variant iodrain -DC2_IO_DRAIN=1 -DC2_SYMW=1               # accumulator drain on the IO warps
ls -la deepsphere-cosmo-tf2_b200/lib/*.so
