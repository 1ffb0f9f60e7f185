#!/bin/bash
# Times one rank's kernels (tools/shard_kernels.py) under a list of variant libraries (python -m simhand_b200.build --variant NAME -D...).
#   gpurun -- bash tools/exp_variants.sh OUTFILE "WORLDS" NAME [NAME ...]       e.g.  ... gpurun_out/x.txt "1 8" m000 r2m000
out=$1; worlds=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  lib=simhand_b200/lib/libsimhand_b200_$v.so
  [ "$v" = product ] && lib=simhand_b200/lib/libsimhand_b200.so
  [ -f $lib ] || { echo "$v : missing" >> $out; continue; }
  for w in $worlds; do
    echo -n "$v (SMH_Q16=${SMH_Q16:-1}) : " >> $out
    SMH_LIB=$PWD/$lib timeout 100 python tools/shard_kernels.py $w 0 2>&1 | tail -1 >> $out
  done
done
cat $out
