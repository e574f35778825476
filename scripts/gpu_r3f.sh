#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -1
for k in 0 1 2; do
  a=$(QLB_LAUNCH_ORDER=$k python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1 | cut -d: -f2)
  b=$(QLB_LAUNCH_ORDER=$k python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2)
  c=$(QLB_LAUNCH_ORDER=$k python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1 | cut -d: -f2)
  d=$(QLB_LAUNCH_ORDER=$k python scripts/shard_time.py benzene_ccpvdz 1 0 2>&1 | tail -1 | cut -d: -f2)
  e=$(QLB_LAUNCH_ORDER=$k python scripts/shard_time.py c40h82_def2svp 8 0 2>&1 | tail -1 | cut -d: -f2)
  echo "order=$k: 1/8 $a | 1/1 $b | c40 $c | benzene $d | c40 1/8 $e"
done
