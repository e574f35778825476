#!/bin/bash
# round-2 GPU session L: bench lines of every workload (N=1), RI workloads, launch list
T=gpurun_out/r2l
timeout 600 python bench.py > ${T}_bench_h2o32.json 2> ${T}_bench_h2o32.err; cut -c1-150 ${T}_bench_h2o32.json
for w in benzene_ccpvdz c40h82_def2svp h2o64_ccpvdz; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 0 > ${T}_bench_$w.json 2> ${T}_bench_$w.err; cut -c1-150 ${T}_bench_$w.json
done
timeout 600 python bench.py --workload h2o32_631gs_rijk --steps 5 --warmup 3 > ${T}_bench_h2o32_rijk.json 2> ${T}_bench_h2o32_rijk.err; cut -c1-150 ${T}_bench_h2o32_rijk.json; tail -2 ${T}_bench_h2o32_rijk.err
timeout 900 python bench.py --workload h2o64_ccpvdz_rijk --steps 3 --warmup 3 > ${T}_bench_h2o64_rijk.json 2> ${T}_bench_h2o64_rijk.err; cut -c1-150 ${T}_bench_h2o64_rijk.json; tail -2 ${T}_bench_h2o64_rijk.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${T}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0 > ${T}_ncu_launches.log 2>&1; tail -1 ${T}_ncu_launches.log | cut -c1-100
