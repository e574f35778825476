#!/bin/bash
# round-2 GPU session T (N GPUs): schedule A/B
N=${1:-2}
T=gpurun_out/r2t_n$N
run() { # name env
env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 4 --workload h2o32_631gs --cpu-seconds 0 > ${T}_$1.json 2> ${T}_$1.err
python -c "
import json; d=json.load(open('${T}_$1.json')); print('$1', d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'rank0 class sum', round(d['roofline']['all_class_kernels']['ms'],2))"
}
run sched X=1
run nosched QLB_NO_SCHEDULE=1
