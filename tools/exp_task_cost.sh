#!/bin/bash
# task-type weights of the plan's CTA cuts (one rank's plan at world 8 / 4 / 1, L2-warm, CUDA events)
out=gpurun_out/r02_task_cost.txt
: > $out
for w in 8 4 1; do
  for tc in "0 0" "0.2 0" "0.4 0" "0.6 0" "0.4 0.4" "0.8 0.4"; do
    set -- $tc
    echo -n "cost transposed=$1 masked=$2 : " >> $out
    SMH_COST_TRANSPOSED=$1 SMH_COST_MASKED=$2 timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
  done
done
cat $out
