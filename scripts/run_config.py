"""Run one BASELINE.json configuration end to end on the GPU: basis -> one-electron matrices -> a few Roothaan steps with
device J/K -> per-class timing of a Fock build; size-independent parity properties are checked (symmetry, linearity,
fused J/K == separate tau=0 ... where affordable).  usage: run_config.py <workload> [nsteps] [tau]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from quantumlab_jl_b200 import libqlb, workloads
from quantumlab_jl_b200.b200shells import B200Shells
from quantumlab_jl_b200.geometry import computeEnergyInteratomicRepulsion
from quantumlab_jl_b200.hartreefock import computeEnergyHartreeFock, evaluateSCFStep

name = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tau = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-12
libqlb.init(0)
geo, shells, nocc = workloads.build(name)
t = time.time()
dev = B200Shells(shells, tau=tau)
t_basis = time.time() - t
N = dev.nbf
t = time.time()
S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
t_1e = time.time() - t
Vnn = computeEnergyInteratomicRepulsion(geo)
out = {"workload": name, "nsh": len(shells), "nbf": N, "natoms": len(geo.atoms), "nocc": nocc, "basis_create_s": t_basis, "one_electron_s": t_1e,
       "S_min_eig": float(np.linalg.eigvalsh(S).min())}
eps, P = evaluateSCFStep(T + V, S, nocc)
libqlb.load().qlb_set_profiling(1)
energies = []
for it in range(nsteps):
    t = time.time()
    J, K = dev.build_jk(P)
    wall = time.time() - t
    st = dev.last_stats
    E = computeEnergyHartreeFock(P, T, V, J, K, Vnn, info=False)
    energies.append(E + Vnn)
    print(f"iter {it}: build wall {wall*1e3:.1f} ms  eri {st['ms_eri']:.1f} ms  kept {st['quartets_kept']}/{st['quartets_unique']}  "
          f"{st['flops_algorithmic']/st['ms_eri']*1e-9:.2f} TF/s(alg)  E_tot {E+Vnn:.8f}", flush=True)
    assert np.array_equal(J, J.T) and np.array_equal(K, K.T)
    eps, P = evaluateSCFStep(T + V + 2 * J - K, S, nocc)
# linearity property on the last density
J1, K1 = dev.build_jk(P, tau=tau)
J2, K2 = dev.build_jk(0.5 * P, tau=tau * 0.5)
out["linearity_max_abs"] = float(max(np.abs(J1 - 2 * J2).max(), np.abs(K1 - 2 * K2).max()))
out["energies"] = energies
out["stats"] = st
print(json.dumps(out))
