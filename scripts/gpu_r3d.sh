#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r3d_tests.log 2>&1; tail -2 gpurun_out/r3d_tests.log
for k in 0 3 6 12 1000; do
  a=$(QLB_DLSMEM_MIN_TILES=$k python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1 | cut -d: -f2)
  b=$(QLB_DLSMEM_MIN_TILES=$k python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2)
  c=$(QLB_DLSMEM_MIN_TILES=$k python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1 | cut -d: -f2)
  d=$(QLB_DLSMEM_MIN_TILES=$k python scripts/shard_time.py benzene_ccpvdz 1 0 2>&1 | tail -1 | cut -d: -f2)
  e=$(QLB_DLSMEM_MIN_TILES=$k python scripts/shard_time.py c40h82_def2svp 8 0 2>&1 | tail -1 | cut -d: -f2)
  echo "min_tiles=$k: 1/8 $a | 1/1 $b | c40 $c | benzene $d | c40 1/8 $e"
done
