#!/bin/bash
T=gpurun_out/r2u
for cfg in "256 8" "256 2" "256 1" "1024 2" "64 8" "1024 8"; do
  set -- $cfg
  echo "target=$1 tiles=$2: $(QLB_YS_TARGET=$1 QLB_YS_TILES=$2 python scripts/shard_time.py h2o32_631gs 8 0 3 2>&1 | tail -2 | cut -d: -f2 | tr '\n' '|')"
done
echo "N=1: $(python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1)"
echo "c40 1/8: $(python scripts/shard_time.py c40h82_def2svp 8 0 2>&1 | tail -1)"
echo "c40 1/1: $(python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1)"
