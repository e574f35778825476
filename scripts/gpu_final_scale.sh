#!/bin/bash
# round-2 final, N GPUs: scaling bench lines (both workload families), RI-JK on 8
N=${1:-2}
T=gpurun_out/final_n$N
python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "multi_gpu or sharded" > ${T}_tests.log 2>&1; tail -1 ${T}_tests.log
run() { # workload steps extra
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $2 --warmup 3 --workload $1 $3 > ${T}_$1.json 2> ${T}_$1.err
python -c "
import json; d=json.load(open('${T}_$1.json')); print('$1', d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
}
run h2o32_631gs 10 "--cpu-seconds 0"
run c40h82_def2svp 10 "--cpu-seconds 0"
if [ "$N" = "8" ] && [ -n "$FULL" ]; then run h2o64_ccpvdz 5 "--cpu-seconds 0"; run h2o64_ccpvdz_rijk 3 ""; fi
