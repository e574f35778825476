#!/bin/bash
# round-2 GPU session E: build5 library (survivor compaction) -- bin width sweep
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200_r2e.so
T=gpurun_out/r2e
python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; tail -6 ${T}_tests.log
for w in 1 8 16 32 64 128; do
  QLB_BIN_W=$w timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32_w$w.log 2>&1; echo "binw=$w: $(grep 'iter 1\|iter 2' ${T}_h2o32_w$w.log | cut -d, -f2 | tr '\n' ' ')"
done
for w in 1 32; do
QLB_BIN_W=$w timeout 200 python scripts/run_config.py benzene_ccpvdz > ${T}_benzene_w$w.log 2>&1; grep "iter 2" ${T}_benzene_w$w.log
QLB_BIN_W=$w timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40_w$w.log 2>&1; grep "iter 2" ${T}_c40_w$w.log
done
