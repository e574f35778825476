#!/bin/bash
# round-2 GPU session N: ncu of the light-flavour kernels
T=gpurun_out/r2n
for k in psss:eri_class_kernelILi1ELi0ELi0ELi0ELi0E ssss:eri_class_kernelILi0ELi0ELi0ELi0ELi0E; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:${k#*:} -s 1 -c 1 \
    -o ${T}_prof_${k%%:*} python scripts/first_run.py h2o32_631gs > ${T}_ncu_${k%%:*}.log 2>&1; tail -1 ${T}_ncu_${k%%:*}.log
done
