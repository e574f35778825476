#!/usr/bin/env python
"""Per-class kernel times (profiling mode: one class at a time) of one rank's share of an nranks-way build on one GPU.
usage: shard_classes.py WORKLOAD NRANKS   (QLB_YS_TARGET etc. from the environment)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("QLB_NO_SCHEDULE", "1")
import numpy as np
import torch

from quantumlab_jl_b200 import libqlb, workloads
from quantumlab_jl_b200.b200shells import B200Shells
from quantumlab_jl_b200.hartreefock import evaluateSCFStep

name, nranks = sys.argv[1], int(sys.argv[2])
lib = libqlb.init(0)
geo, shells, nocc = workloads.build(name)
dev = B200Shells(shells, tau=1e-12)
N = dev.nbf
S = dev.one_electron("overlap")
H = np.asfortranarray(dev.one_electron("kinetic") + dev.one_electron("nuclear", geo))
_, P = evaluateSCFStep(H, S, nocc)
F = dev.fock(H, P)
_, P = evaluateSCFStep(F, S, nocc)
dP = torch.from_numpy(np.ascontiguousarray(np.asarray(P).T)).cuda()
dJK = torch.empty(2 * N * N + 64, dtype=torch.float64, device="cuda")
pJ, pK = C.c_void_p(dJK.data_ptr()), C.c_void_p(dJK.data_ptr() + 8 * N * N)
libqlb.check(lib.qlb_set_profiling(1))
acc = None
for it in range(5):
    st = libqlb.Stats()
    libqlb.check(lib.qlb_build_jk_device(dev.handle, C.c_void_p(dP.data_ptr()), 1e-12, pJ, pK, 0, nranks, None, C.byref(st)))
    torch.cuda.synchronize()
    d = st.as_dict()["ms_class"]
    if it >= 2:
        acc = d if acc is None else {k: min(acc[k], v) for k, v in d.items()}
print(name, nranks, os.environ.get("QLB_TPW", os.environ.get("QLB_YS_TARGET", "-")), " ".join(f"{k}={v:.3f}" for k, v in acc.items()), "sum=%.3f" % sum(acc.values()), flush=True)
