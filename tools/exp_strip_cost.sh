#!/bin/bash
# one rank's kernels (plan of rank 0 at world 8 / 4 / 2 / 1, L2-warm, CUDA events) for the strip-cost weights of the two cuts
out=gpurun_out/r02_strip_cost4.txt
: > $out
for w in 8 4 2 1; do
  for sc in "0.5 4" "0 4" "1 3" "0.5 6"; do
    set -- $sc
    echo -n "strip_cost fwd=$1 bwd=$2 : " >> $out
    SMH_STRIP_COST_FWD=$1 SMH_STRIP_COST=$2 timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
  done
done
cat $out
