#!/bin/bash
# round-2 GPU session G: build7 library (packed pair records, density-bound rows in shared memory)
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200_r2g.so
T=gpurun_out/r2g
python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; tail -4 ${T}_tests.log
timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32.log 2>&1; grep -E "iter" ${T}_h2o32.log | cut -d, -f2-4
timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40.log 2>&1; grep "iter" ${T}_c40.log
