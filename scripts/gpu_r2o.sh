#!/bin/bash
# round-2 GPU session O: 128-wide DMMA GEMM tiles -- tests, RI bench A/B, ncu of the dgemm kernels
T=gpurun_out/r2o
python -m pytest tests/test_gpu_configs.py tests/test_ri.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -4 ${T}_tests.log
timeout 600 python bench.py --workload h2o32_631gs_rijk --steps 5 --warmup 3 > ${T}_h2o32_rijk.json 2> ${T}_h2o32_rijk.err; python -c "
import json; d=json.load(open('${T}_h2o32_rijk.json')); print('big', d['ms_per_step'], d['roofline']['frac'])"
QLB_DGEMM_SMALL=1 timeout 600 python bench.py --workload h2o32_631gs_rijk --steps 5 --warmup 3 > ${T}_h2o32_rijk_small.json 2> ${T}_h2o32_rijk_small.err; python -c "
import json; d=json.load(open('${T}_h2o32_rijk_small.json')); print('small', d['ms_per_step'], d['roofline']['frac'])"
timeout 900 python bench.py --workload h2o64_ccpvdz_rijk --steps 3 --warmup 2 > ${T}_h2o64_rijk.json 2> ${T}_h2o64_rijk.err; python -c "
import json; d=json.load(open('${T}_h2o64_rijk.json')); print('big h2o64', d['ms_per_step'], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm -s 8 -c 4 -o ${T}_prof_dgemm python bench.py --workload h2o32_631gs_rijk --steps 1 --warmup 1 > ${T}_ncu_dgemm.log 2>&1; tail -1 ${T}_ncu_dgemm.log
