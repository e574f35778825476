#!/bin/bash
# round-2 GPU session F: build6 library (exclusive dispatch chains, compaction inlined, exact bound order)
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200_r2f.so
T=gpurun_out/r2f
python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; tail -4 ${T}_tests.log
timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32.log 2>&1; grep -E "iter" ${T}_h2o32.log | cut -d, -f2-4
timeout 200 python scripts/run_config.py benzene_ccpvdz > ${T}_benzene.log 2>&1; grep "iter" ${T}_benzene.log
timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40.log 2>&1; grep "iter" ${T}_c40.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > ${T}_bench.json 2> ${T}_bench.err; cut -c1-200 ${T}_bench.json
for k in psss:eri_class_kernelILi1ELi0ELi0ELi0ELi0E dpdp:eri_coop_kernelILi2ELi1ELi2ELi1ELi0E dddd:eri_coop_kernelILi2ELi2ELi2ELi2ELi0E; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:${k#*:} -s 1 -c 1 \
    -o ${T}_prof_${k%%:*} python scripts/first_run.py h2o32_631gs > ${T}_ncu_${k%%:*}.log 2>&1; tail -1 ${T}_ncu_${k%%:*}.log
done
