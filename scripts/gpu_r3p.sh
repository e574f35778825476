#!/bin/bash
for v in _base ""; do
  export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200$v.so
  echo "lib$v: $(python scripts/shard_classes.py h2o32_631gs 1 2>&1 | tail -1 | cut -d' ' -f10) c40 $(python scripts/shard_classes.py c40h82_def2svp 1 2>&1 | tail -1 | cut -d' ' -f10) h2o64 $(python scripts/shard_classes.py h2o64_ccpvdz 1 2>&1 | tail -1 | cut -d' ' -f10)"
done
export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200.so
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -1
