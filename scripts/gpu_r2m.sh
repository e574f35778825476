#!/bin/bash
# round-2 GPU session M: light flavour (bra primitives in shared memory, point-charge branch) for the L <= 2 classes
T=gpurun_out/r2m
python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; tail -4 ${T}_tests.log
timeout 200 python scripts/first_run.py h2o32_631gs > ${T}_h2o32.log 2>&1; grep -E "iter" ${T}_h2o32.log | cut -d, -f2-4
python - <<'PY'
import json,re
s=open('gpurun_out/r2m_h2o32.log').read()
i=s.rindex('"ms_class"'); print(s[i:i+420].replace('\n',' '))
PY
timeout 200 python scripts/run_config.py benzene_ccpvdz > ${T}_benzene.log 2>&1; grep "iter" ${T}_benzene.log
timeout 300 python scripts/run_config.py c40h82_def2svp > ${T}_c40.log 2>&1; grep "iter" ${T}_c40.log
