#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r3s_tests.log 2>&1; tail -1 gpurun_out/r3s_tests.log
echo "$(python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1 | cut -d: -f2) | c40 $(python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1 | cut -d: -f2) | benzene $(python scripts/shard_time.py benzene_ccpvdz 1 0 2>&1 | tail -1 | cut -d: -f2) | 1/8 $(python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1 | cut -d: -f2)"
python scripts/shard_classes.py h2o32_631gs 1 2>&1 | tail -1
