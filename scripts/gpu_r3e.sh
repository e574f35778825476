#!/bin/bash
for k in 3 7 15 23 39; do
  a=$(QLB_NSIDE=$k python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1 | cut -d: -f2)
  b=$(QLB_NSIDE=$k python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2)
  d=$(QLB_NSIDE=$k python scripts/shard_time.py benzene_ccpvdz 1 0 2>&1 | tail -1 | cut -d: -f2)
  e=$(QLB_NSIDE=$k python scripts/shard_time.py c40h82_def2svp 8 0 2>&1 | tail -1 | cut -d: -f2)
  echo "nside=$k: 1/8 $a | 1/1 $b | benzene $d | c40 1/8 $e"
done
