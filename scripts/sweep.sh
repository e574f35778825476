#!/bin/bash
for cfg in "256 8" "256 16" "128 16" "512 4" "64 32" "1024 8" "256 2"; do
  set -- $cfg
  QLB_YS_TARGET=$1 QLB_YS_TILES=$2 timeout 250 python scripts/first_run.py h2o32_631gs > gpurun_out/sweep_$1_$2.log 2>&1
  echo "target=$1 tiles=$2: $(grep 'iter 2' gpurun_out/sweep_$1_$2.log | cut -d, -f3)"
done
