#!/bin/bash
# compute-sanitizer over every kernel family (tools/sanitize_target.py): memcheck, racecheck (shared-memory hazards,
# incl. the mbarrier-ordered TMA / tcgen05 pipelines), synccheck (barrier misuse).  Logs -> gpurun_out/r02_sanitizer_*.log,
# summaries are copied to profiles/.     gpurun -- bash tools/sanitize.sh
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  for what in step shard head; do
    log=gpurun_out/r02_sanitizer_${tool}_${what}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $what > $log 2>&1
    echo "== $tool $what: exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done | tee gpurun_out/r02_sanitizer_summary.txt
