#!/bin/bash
# mpjpe modes and strip-cost experiments (one rank's plan, L2-warm, CUDA events)
out=gpurun_out/r02_mpjpe_modes.txt
: > $out
for w in 1 8 2; do
  for q in 1 0; do
    for mode in 1s 2s 4s 1d 2d 4d; do
      echo -n "SMH_Q16=$q mode=$mode : " >> $out
      SMH_Q16=$q SMH_MPJPE_MODE=$mode timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
    done
  done
done
out2=gpurun_out/r02_strip_cost.txt
: > $out2
for w in 1 8 2 4; do
  for sc in 0 1 2 4; do
    echo -n "strip_cost=$sc : " >> $out2
    SMH_STRIP_COST=$sc timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out2
  done
done
cat $out $out2
