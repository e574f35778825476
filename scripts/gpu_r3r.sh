#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r3r_tests.log 2>&1; tail -1 gpurun_out/r3r_tests.log
for v in _base ""; do
  export QLB200_LIB=$PWD/quantumlab_jl_b200/libqlb200$v.so
  echo "lib$v: $(python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2) | c40 $(python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1 | cut -d: -f2)"
  python scripts/shard_classes.py h2o32_631gs 1 2>&1 | tail -1
done
