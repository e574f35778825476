#!/bin/bash
# round-2 GPU session C: build3 library (row walk + device density bound + bin/index pair order + coop v1)
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200_r2c.so
T=gpurun_out/r2c
python -m pytest tests -m gpu -x -q -k "not jkfit" > ${T}_tests.log 2>&1; tail -5 ${T}_tests.log
for w in 8 16 32 64; do
  QLB_BIN_W=$w timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32_w$w.log 2>&1; echo "binw=$w: $(grep 'iter 1\|iter 2' ${T}_h2o32_w$w.log | cut -d, -f2-3 | tr '\n' ' ')"
done
timeout 200 python scripts/run_config.py benzene_ccpvdz > ${T}_benzene.log 2>&1; grep iter ${T}_benzene.log
timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40.log 2>&1; grep iter ${T}_c40.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > ${T}_bench.json 2> ${T}_bench.err; cat ${T}_bench.json | cut -c1-400
# ncu: the two dominant light classes, second build of the run
for k in psss:ILi1ELi0ELi0ELi0ELi0E ssss:ILi0ELi0ELi0ELi0ELi0E; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:eri_class_kernel${k#*:} -s 1 -c 1 \
    -o ${T}_prof_${k%%:*} python scripts/first_run.py h2o32_631gs > ${T}_ncu_${k%%:*}.log 2>&1; tail -2 ${T}_ncu_${k%%:*}.log
done
