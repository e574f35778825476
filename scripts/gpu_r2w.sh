#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; tail -2 gpurun_out/r2w_tests.log
for cfg in "8 8" "16 8" "32 8" "64 8" "256 8"; do
  set -- $cfg
  a=$(QLB_YS_TARGET=$1 QLB_YS_TILES=$2 python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1 | cut -d: -f2)
  b=$(QLB_YS_TARGET=$1 QLB_YS_TILES=$2 python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2)
  c=$(QLB_YS_TARGET=$1 QLB_YS_TILES=$2 python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1 | cut -d: -f2)
  echo "target=$1 tiles=$2: 1/8 $a | 1/1 $b | c40 $c"
done
