#!/bin/bash
# compute-sanitizer over every kernel family (tools/sanitize_target.py): memcheck on all three workloads, synccheck (barrier
# misuse) and racecheck (shared-memory hazards) on the fused step and the projection head.  Logs -> gpurun_out/r02_sanitizer_*.log;
# the summary is copied to profiles/.     gpurun -- bash tools/sanitize.sh
set -u
mkdir -p gpurun_out
# SAN_JOBS="tool:workload ..." and SAN_TIMEOUT (seconds per job) restrict the run when GPU time is short
JOBS=${SAN_JOBS:-"memcheck:step memcheck:shard memcheck:head synccheck:step synccheck:head racecheck:head racecheck:step"}
for job in $JOBS; do
  set -- ${job/:/ }
  log=gpurun_out/r02_sanitizer_$1_$2.log
  timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $1 --print-limit 10 python tools/sanitize_target.py $2 > $log 2>&1
  echo "== $1 $2: exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
done | tee gpurun_out/r02_sanitizer_summary.txt
