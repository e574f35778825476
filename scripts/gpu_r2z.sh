#!/bin/bash
T=gpurun_out/r2z
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:eri_class_kernelILi0ELi0ELi0ELi0ELi0E -s 3 -c 1 -o ${T}_prof_ssss_8th python scripts/shard_time.py h2o32_631gs 8 0 > ${T}_ncu1.log 2>&1; tail -1 ${T}_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:eri_class_kernelILi1ELi1ELi1ELi1ELi0E -s 3 -c 1 -o ${T}_prof_pppp_8th python scripts/shard_time.py h2o32_631gs 8 0 > ${T}_ncu2.log 2>&1; tail -1 ${T}_ncu2.log
