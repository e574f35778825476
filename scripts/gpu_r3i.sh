#!/bin/bash
# field-major pair_info layout: correctness + timing
python -m pytest tests -m gpu -q -x > gpurun_out/r3i_tests.log 2>&1; tail -2 gpurun_out/r3i_tests.log
python scripts/shard_time.py h2o32_631gs 1 0 2>&1 | tail -1
python scripts/shard_time.py h2o32_631gs 8 0 2>&1 | tail -1
python scripts/shard_time.py c40h82_def2svp 1 0 2>&1 | tail -1
python scripts/shard_time.py benzene_ccpvdz 1 0 2>&1 | tail -1
python scripts/shard_classes.py h2o32_631gs 1 2>&1 | tail -1
