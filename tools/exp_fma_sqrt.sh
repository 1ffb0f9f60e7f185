#!/bin/bash
# MPJPE pipe balance: how many of the 21 square roots per pair go through the FMA pipe (sqrt2_fma_pipe) instead of MUFU.SQRT.
# Variant libraries:  for v in 000 040 440 042 442 242 24a; do python -m simhand_b200.build --variant m$v -DSMH_MPJPE_FMA_SQRT_MASK=0x$v; done
#                     python -m simhand_b200.build --variant x3 -DSMH_MPJPE_EXACT_CTAS=3      (exact form at 3 CTAs per SM)
# One rank's kernels on one GPU, CUDA events (tools/shard_kernels.py).   gpurun -- bash tools/exp_fma_sqrt.sh
out=gpurun_out/r02_mpjpe_fma_sqrt.txt
mkdir -p gpurun_out
: > $out
for v in 000 040 440 042 442 242 24a; do
  lib=simhand_b200/lib/libsimhand_b200_m$v.so
  [ -f $lib ] || continue
  for w in 1 8; do
    echo -n "mask=0x$v : " >> $out
    SMH_LIB=$PWD/$lib SMH_Q16=1 timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
  done
done
for lib in libsimhand_b200.so libsimhand_b200_x3.so; do
  [ -f simhand_b200/lib/$lib ] || continue
  for w in 1 8; do
    echo -n "exact form, $lib : " >> $out
    SMH_LIB=$PWD/simhand_b200/lib/$lib SMH_Q16=0 timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
  done
done
cat $out
