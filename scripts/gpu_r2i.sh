#!/bin/bash
# round-2 GPU session I: grid-shape sweep on the build8 library (longer screening walks per warp -> fuller bucketed batches)
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200_r2h.so
export QLB_NO_DLSMEM=1
T=gpurun_out/r2i
for cfg in "256 8" "64 8" "32 16" "16 32" "8 64" "4 128"; do
  set -- $cfg
  QLB_YS_TARGET=$1 QLB_YS_TILES=$2 timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32_$1_$2.log 2>&1
  echo "target=$1 tiles=$2: $(grep 'iter 1\|iter 2' ${T}_h2o32_$1_$2.log | cut -d, -f2 | tr '\n' ' ')"
done
for cfg in "256 8" "16 32"; do
  set -- $cfg
  QLB_YS_TARGET=$1 QLB_YS_TILES=$2 timeout 200 python scripts/run_config.py benzene_ccpvdz > ${T}_benzene_$1_$2.log 2>&1; grep "iter 2" ${T}_benzene_$1_$2.log
  QLB_YS_TARGET=$1 QLB_YS_TILES=$2 timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40_$1_$2.log 2>&1; grep "iter 2" ${T}_c40_$1_$2.log
done
