"""RI-J / RI-K demonstration on a water cluster with the synthetic even-tempered auxiliary basis: device timings of
the dense FP64 contractions and the fitting error against the conventional (exact) J/K of the same density."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from quantumlab_jl_b200 import libqlb, workloads
from quantumlab_jl_b200.b200shells import B200Shells
from quantumlab_jl_b200.hartreefock import evaluateSCFStep
from quantumlab_jl_b200.resolutionidentity import B200RI

name = sys.argv[1] if len(sys.argv) > 1 else "h2o4_631gs"
libqlb.init(0)
geo, shells, nocc = workloads.build(name)
aux = workloads.even_tempered_aux(geo)
dev = B200Shells(shells)
S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
_, P = evaluateSCFStep(T + V, S, nocc)
F = dev.fock(T + V, P)
_, P = evaluateSCFStep(F, S, nocc)
J, K = dev.build_jk(P)
t = time.time()
ri = B200RI(dev, aux)
t_create = time.time() - t
for _ in range(3):
    Jr, Kr = ri.build_jk(P)
N, Nx = ri.nbf, ri.naux
flops = 2.0 * 2 * N * N * Nx + 2.0 * 2 * N ** 3 * Nx
out = {"workload": name, "N": N, "Naux": Nx, "ri_create_s": t_create, "ri_jk_ms": ri.last_ms, "conventional_jk_ms": dev.last_stats["ms_total"],
       "ri_contraction_tflops": flops / (ri.last_ms * 1e-3) * 1e-12,
       "max_abs_J_err": float(np.abs(Jr - J).max()), "max_abs_K_err": float(np.abs(Kr - K).max()),
       "rel_EJ_err": float(abs(np.trace(Jr @ P) - np.trace(J @ P)) / abs(np.trace(J @ P))),
       "rel_EK_err": float(abs(np.trace(Kr @ P) - np.trace(K @ P)) / abs(np.trace(K @ P)))}
print(json.dumps(out))
