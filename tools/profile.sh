#!/bin/bash
# ncu evidence for one step at the graded size (run on the GPU box via gpurun; outputs under gpurun_out/).
#   1. launch list with device time per launch (cold-cache, serialised: compare shares)
#   2. full-section capture of the three hot kernels (mpjpe, forward sweep, backward sweep)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mpjpe_kernel|sweep_tc_kernel" -s 6 -c 3 \
    -o gpurun_out/prof -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
