#!/bin/bash
for t in 4 8 16 32 64 128 256 512; do
  QLB_YS_TARGET=$t python scripts/shard_classes.py h2o32_631gs 8 2>&1 | tail -1
done
for t in 8 16 32 64 128; do
  QLB_YS_TARGET=$t python scripts/shard_classes.py h2o32_631gs 1 2>&1 | tail -1
done
for t in 8 16 32 64 128; do
  QLB_YS_TARGET=$t python scripts/shard_classes.py c40h82_def2svp 1 2>&1 | tail -1
done
