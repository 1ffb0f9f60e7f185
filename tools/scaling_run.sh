#!/bin/bash
# Round-end style scaling measurement on one 8-GPU box (run via `gpurun --gpus 8`): bench.py at N = 1, 2, 4, 8 back to
# back, the sharded parity checks, the per-stage breakdown at N = 8 and the end-to-end example.  Outputs: gpurun_out/.
mkdir -p gpurun_out
run() { # N port
  if [ "$1" = 1 ]; then timeout 300 python bench.py --gpus 1 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" \
       bench.py --gpus "$1" > gpurun_out/scale_n$1.json 2> gpurun_out/scale_n$1.err; fi
  echo "N=$1 rc=$?"; cut -c1-260 gpurun_out/scale_n$1.json
}
run 1 0; run 2 29701; run 4 29702; run 8 29703
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29704 \
    tools/dist_check.py 8192 > gpurun_out/dist_check_n8.log 2>&1; grep -E "rank [0-9]/8" gpurun_out/dist_check_n8.log | head -16
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29705 \
    tools/dist_profile.py > gpurun_out/dist_profile_n8.log 2>&1; grep -E "^rank 0" gpurun_out/dist_profile_n8.log
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/dist_tests_n8.log 2>&1; tail -3 gpurun_out/dist_tests_n8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29706 \
    examples/e2e_step.py --batch 8192 --steps 5 > gpurun_out/e2e_n8.log 2>&1; grep -E "^\{" gpurun_out/e2e_n8.log | cut -c1-400
