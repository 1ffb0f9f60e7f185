#!/bin/bash
# One 8-GPU box: sharded parity at full size, phase clocks, bench at 8 / 4 / 2 GPUs (fused exchange; the unfused one at 8 for
# comparison).   gpurun --gpus 8 -- bash tools/scaling_run.sh
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
mkdir -p gpurun_out
run 300 8 29721 tools/dist_check.py 8192 fused 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/r02_dist_check_n8.log
run 200 8 29722 tools/shard_phase_times.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | sed "s/\\\\n/\n/g" | tee gpurun_out/r02_phase_clocks_n8.txt
for cfg in "8 fused" "8 peer" "4 fused" "2 fused"; do
  set -- $cfg
  run 300 $1 29723 bench.py --gpus $1 --steps 300 --warmup 5 --transport $2 > gpurun_out/r02_bench_n$1_$2.json 2> gpurun_out/r02_bench_n$1_$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_n$1_$2.json"))
    print("n$1 $2", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "exact", round(d["exact_weights"]["value"], 1), d.get("kernels_ms"), d.get("clocks"))
except Exception as e:
    print("n$1 $2 FAILED", e)
PY
done
timeout 300 python bench.py --steps 300 > gpurun_out/r02_bench_n1_box8.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_box8.json')); print('n1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'exact', round(d['exact_weights']['value'],1))"
