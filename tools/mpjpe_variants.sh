#!/bin/bash
# times the MPJPE kernel variants (SMH_MPJPE_VARIANT tuning knob) at the graded size
for v in 0 1 2 3 4; do
  SMH_MPJPE_VARIANT=$v timeout 120 python bench.py --steps 40 --warmup 5 --no-graph > gpurun_out/mv_$v.json 2> gpurun_out/mv_$v.err
  python -c "import json; d=json.load(open('gpurun_out/mv_$v.json')); print('variant $v mpjpe_ms', round(d['kernels_ms']['mpjpe_kernel'],4), 'step', round(d['ms_per_step'],4))" || tail -3 gpurun_out/mv_$v.err
done
