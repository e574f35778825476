#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r3q_tests.log 2>&1; tail -1 gpurun_out/r3q_tests.log
for k in 0 1; do
  a=$(QLB_SPLIT_STREAMS=$k python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1 | cut -d: -f2)
  b=$(QLB_SPLIT_STREAMS=$k python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2)
  d=$(QLB_SPLIT_STREAMS=$k python scripts/shard_time.py benzene_ccpvdz 1 0 2>&1 | tail -1 | cut -d: -f2)
  e=$(QLB_SPLIT_STREAMS=$k python scripts/shard_time.py c40h82_def2svp 8 0 2>&1 | tail -1 | cut -d: -f2)
  echo "split_streams=$k: 1/8 $a | 1/1 $b | benzene $d | c40 1/8 $e"
done
