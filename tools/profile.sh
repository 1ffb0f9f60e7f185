#!/bin/bash
# ncu evidence for one step at the graded size (run on the GPU box via gpurun; outputs under gpurun_out/).
#   1. launch list with device time per launch (cold-cache, serialised: compare shares)
#   2. full-section capture of the three hot kernels (mpjpe, forward sweep, backward sweep), relaxed-weights (default) and
#      exact-weights mode, and of the projection head's kernels
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/r02_ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mpjpe_kernel|sweep_tc_kernel" -s 3 -c 3 \
    -o gpurun_out/r02_prof -f python bench.py --steps 2 --warmup 3 > gpurun_out/r02_ncu_full_bench.log 2>&1
# the exact-weights step of the same command: its kernels are the Q16 = false instantiations (matched on the demangled name)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"mpjpe_kernel<(\(bool\))?0|sweep_tc_kernel<(\(bool\))?[01], (\(bool\))?1, (\(bool\))?0" -c 3 \
    -o gpurun_out/r02_prof_exact -f python bench.py --steps 2 --warmup 3 > gpurun_out/r02_ncu_full_bench_exact.log 2>&1
[ -n "$SKIP_HEAD" ] || ncu --set full --clock-control none --import-source on -k regex:"head_" -s 8 -c 4 \
    -o gpurun_out/r02_prof_head -f python tools/bench_head.py > gpurun_out/r02_ncu_full_head.log 2>&1
ls -la gpurun_out | tail -20
