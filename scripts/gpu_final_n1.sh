#!/bin/bash
# round-2 final, 1 GPU: the whole GPU test suite, smoke(), the bench line of every workload, both arms
T=gpurun_out/final
python -m pytest tests -m gpu -q > ${T}_tests.log 2>&1; tail -3 ${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 900 python bench.py > ${T}_bench_h2o32.json 2> ${T}_bench_h2o32.err; cut -c1-160 ${T}_bench_h2o32.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${T}_bench_reference.json 2> ${T}_bench_reference.err; cut -c1-200 ${T}_bench_reference.json
for w in benzene_ccpvdz c40h82_def2svp h2o64_ccpvdz; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 0 > ${T}_bench_$w.json 2> ${T}_bench_$w.err; cut -c1-160 ${T}_bench_$w.json
done
timeout 600 python bench.py --workload h2o32_631gs_rijk --steps 5 --warmup 3 > ${T}_bench_h2o32_rijk.json 2> ${T}_bench_h2o32_rijk.err; cut -c1-160 ${T}_bench_h2o32_rijk.json
timeout 900 python bench.py --workload h2o64_ccpvdz_rijk --steps 3 --warmup 3 > ${T}_bench_h2o64_rijk.json 2> ${T}_bench_h2o64_rijk.err; cut -c1-160 ${T}_bench_h2o64_rijk.json
