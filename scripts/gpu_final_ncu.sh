#!/bin/bash
# round-2 final captures (after bench.py / the tests have exited 0 without ncu): --set full of the dominant light kernels, the
# cooperative high-L kernels, the DMMA GEMM and the small HBM-bound kernels; the launch list of the bench command
T=gpurun_out/final
W="python scripts/first_run.py h2o32_631gs"
cap() { # label regex skip
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$2 -s $3 -c 1 -o ${T}_prof_$1 $W > ${T}_ncu_$1.log 2>&1; tail -1 ${T}_ncu_$1.log
}
cap psss eri_class_kernelILi1ELi0ELi0ELi0ELi0E 1
cap ssss eri_class_kernelILi0ELi0ELi0ELi0ELi0E 1
cap dsps eri_class_kernelILi2ELi0ELi1ELi0ELi0E 1
cap dppp eri_coop_kernelILi2ELi1ELi1ELi1E 3
cap dpdp eri_coop_kernelILi2ELi1ELi2ELi1E 3
cap dddd eri_coop_kernelILi2ELi2ELi2ELi2E 6
timeout 500 ncu --set full --clock-control none --kernel-name-base mangled -k 'regex:pair_prim_kernel|density_block_max|pair_dl_kernel|symmetrize_kernel|fock_kernel' -c 8 -o ${T}_prof_small $W > ${T}_ncu_small.log 2>&1; tail -1 ${T}_ncu_small.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm -s 8 -c 2 -o ${T}_prof_dgemm python bench.py --workload h2o32_631gs_rijk --steps 1 --warmup 1 > ${T}_ncu_dgemm.log 2>&1; tail -1 ${T}_ncu_dgemm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file ${T}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0 > ${T}_ncu_launches.log 2>&1; tail -1 ${T}_ncu_launches.log | cut -c1-100
