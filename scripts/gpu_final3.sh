#!/bin/bash
T=gpurun_out/final
timeout 600 python bench.py > ${T}_bench_h2o32.json 2> ${T}_bench_h2o32.err; cut -c1-160 ${T}_bench_h2o32.json
for w in c40h82_def2svp benzene_ccpvdz h2o64_ccpvdz; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 0 > ${T}_bench_$w.json 2> ${T}_bench_$w.err; cut -c1-160 ${T}_bench_$w.json
done
