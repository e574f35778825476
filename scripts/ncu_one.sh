#!/bin/bash
# usage: ncu_one.sh <tag> <skip> -- full ncu capture of ONE class kernel launch (the <skip>-th eri_class_kernel launch)
ncu --set full --clock-control none --import-source on -k regex:eri_class_kernel -s $2 -c 1 -o gpurun_out/prof_$1 \
    python bench.py --steps 2 --warmup 1 --cpu-seconds 0 > gpurun_out/ncu_$1.log 2>&1
ls -la gpurun_out/prof_$1.ncu-rep
