"""Exploratory run: per-class timing of one Fock build on a named workload (not the bench contract)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from quantumlab_jl_b200 import libqlb, workloads
from quantumlab_jl_b200.b200shells import B200Shells
from quantumlab_jl_b200.hartreefock import evaluateSCFStep

name = sys.argv[1] if len(sys.argv) > 1 else "h2o32_631gs"
tau = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
libqlb.init(0)
print(libqlb.device_info(), "fp64 probe TF/s:", libqlb.fp64_peak_probe())
geo, shells, nocc = workloads.build(name)
t = time.time()
dev = B200Shells(shells)
print("basis create s:", time.time() - t, "nsh", len(shells), "N", dev.nbf)
t = time.time()
S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
print("1e s:", time.time() - t)
eps, P = evaluateSCFStep(T + V, S, nocc)
libqlb.load().qlb_set_profiling(1)
for it in range(3):
    t = time.time()
    J, K = dev.build_jk(P, tau=tau)
    wall = time.time() - t
    st = dev.last_stats
    print(f"iter {it}: wall {wall*1e3:.1f} ms, device total {st['ms_total']:.1f} ms, eri {st['ms_eri']:.1f} ms, kept {st['quartets_kept']}/{st['quartets_unique']}"
          f" pairs {st['pairs_significant']}/{st['pairs_total']} flops {st['flops_algorithmic']:.3e} -> {st['flops_algorithmic']/st['ms_eri']*1e-9:.3f} TF/s")
    F = T + V + 2 * J - K
    eps, P = evaluateSCFStep(F, S, nocc)
print(json.dumps(st, indent=1))
E = 2 * np.trace((T + V) @ P)
print("E1", E)
