#!/bin/bash
# round-2 GPU session P: DMMA GEMM variants (split-K, hoisted packed-symmetric index math, tile-size dispatch) on the RI workloads
T=gpurun_out/r2p
python -m pytest tests/test_gpu_configs.py tests/test_ri.py -m gpu -q -x > ${T}_tests.log 2>&1; tail -2 ${T}_tests.log
run() { # name env workload steps
  env $2 timeout 900 python bench.py --workload $3 --steps $4 --warmup 2 > ${T}_$1.json 2> ${T}_$1.err; python -c "
import json; d=json.load(open('${T}_$1.json')); print('$1', round(d['ms_per_step'],2), round(d['roofline']['frac'],4))"
}
run h2o32_default X=1 h2o32_631gs_rijk 5
run h2o32_small QLB_DGEMM_SMALL=1 h2o32_631gs_rijk 5
run h2o32_big QLB_DGEMM_BIG=1 h2o32_631gs_rijk 5
run h2o64_default X=1 h2o64_ccpvdz_rijk 3
run h2o64_small QLB_DGEMM_SMALL=1 h2o64_ccpvdz_rijk 3
