#!/usr/bin/env python
"""Time ONE rank's share of an nranks-way sharded build on a single GPU (no NCCL): qlb_build_jk_device(rank, nranks) with the
round-robin row deal (QLB_NO_SCHEDULE).  Used to tune the small-share regime (N = 8) without paying for 8 GPUs.
usage: shard_time.py WORKLOAD NRANKS [RANKS...]"""
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
ranks = [int(x) for x in sys.argv[3:]] or [0]
lib = libqlb.init(0)
geo, shells, nocc = workloads.build(name)
dev = B200Shells(shells, tau=1e-12)
N = dev.nbf
S = dev.one_electron("overlap")
H = np.asfortranarray(dev.one_electron("kinetic") + dev.one_electron("nuclear", geo))
_, P = evaluateSCFStep(H, S, nocc)
F = dev.fock(H, P)
_, P = evaluateSCFStep(F, S, nocc)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)
dP = torch.from_numpy(np.ascontiguousarray(np.asarray(P).T)).cuda()
dJK = torch.empty(2 * N * N + 64, dtype=torch.float64, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
pJ, pK = C.c_void_p(dJK.data_ptr()), C.c_void_p(dJK.data_ptr() + 8 * N * N)
for r in ranks:
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        libqlb.check(lib.qlb_build_jk_device(dev.handle, C.c_void_p(dP.data_ptr()), 1e-12, pJ, pK, r, nranks, sptr, None))
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name} rank {r}/{nranks}: ms " + " ".join(f"{t:.3f}" for t in ts[3:]), flush=True)
