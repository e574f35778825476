#!/bin/bash
# round-2 GPU session H: build8 library (coop v3, bucketed queues, packed pair records, smem density rows, cheap large-T Boys)
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200_r2h.so
T=gpurun_out/r2h
python -m pytest tests/test_golden.py -m gpu -q -x > ${T}_tests0.log 2>&1; tail -3 ${T}_tests0.log
QLB_NO_DLSMEM=1 python -m pytest tests/test_golden.py -m gpu -q -x > ${T}_tests0b.log 2>&1; tail -3 ${T}_tests0b.log
python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; tail -4 ${T}_tests.log
timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32.log 2>&1; grep -E "iter" ${T}_h2o32.log | cut -d, -f2-4
QLB_NO_DLSMEM=1 timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32_nodl.log 2>&1; grep -E "iter" ${T}_h2o32_nodl.log | cut -d, -f2-4
timeout 200 python scripts/run_config.py benzene_ccpvdz > ${T}_benzene.log 2>&1; grep "iter" ${T}_benzene.log
timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40.log 2>&1; grep "iter" ${T}_c40.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 0 > ${T}_bench.json 2> ${T}_bench.err; cut -c1-200 ${T}_bench.json
