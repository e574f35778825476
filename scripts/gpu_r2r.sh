#!/bin/bash
# round-2 GPU session R (N GPUs): multi-GPU tests + scaling bench lines
N=${1:-2}
T=gpurun_out/r2r_n$N
python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "multi_gpu or sharded" > ${T}_tests.log 2>&1; tail -2 ${T}_tests.log
for w in h2o32_631gs c40h82_def2svp; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $w --cpu-seconds 0 > ${T}_$w.json 2> ${T}_$w.err
python -c "
import json; d=json.load(open('${T}_$w.json')); print('$w', d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done
