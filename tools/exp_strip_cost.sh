#!/bin/bash
# strip-cost weights of the plan's CTA cuts (one rank's plan, L2-warm, CUDA events)
out=gpurun_out/r02_strip_cost3.txt
: > $out
for w in 8 4 2 1; do
  for sc in 0 1 2 4 6; do
    echo -n "strip_cost=$sc : " >> $out
    SMH_STRIP_COST=$sc timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
  done
done
cat $out
