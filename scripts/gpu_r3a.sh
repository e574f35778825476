#!/bin/bash
# banded grids: correctness + per-class sweep of tiles per walk
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/r3a_tests.log 2>&1; tail -2 gpurun_out/r3a_tests.log
for t in 1 2 4 8 16 32; do
  QLB_TPW=$t python scripts/shard_classes.py h2o32_631gs 8 2>&1 | tail -1
done
for t in 1 2 4 8 16 32; do
  QLB_TPW=$t python scripts/shard_classes.py h2o32_631gs 1 2>&1 | tail -1
done
for t in 2 4 8 16 32; do
  QLB_TPW=$t python scripts/shard_classes.py c40h82_def2svp 1 2>&1 | tail -1
done
for t in 2 4 8 16; do
  echo "TPW=$t $(QLB_TPW=$t python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1) $(QLB_TPW=$t python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1)"
done
