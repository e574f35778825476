"""Golden J/K fixtures for the BASELINE.json configurations at FULL size, from the CPU oracle's MD path
(oracle/ql_oracle.c: 8-fold symmetry, the integer Schwarz x density rule, tau = 1e-12), plus the energies of the first
Roothaan steps of an oracle-driven SCF.  The density of the J/K fixtures is a seeded, spatially decaying symmetric
matrix (no SCF needed).

Stored per configuration (tests/golden/<name>_jk.npz):
  cols, Jc, Kc      every `stride`-th column of J and K
  Jr, Kr            J @ r and K @ r for a seeded vector r: a checksum that every element of J and K enters
  sumJ, sumK, trJP, trKP, counts = (kept, unique) with the ORACLE's Schwarz bounds
The GPU test compares all of it, and compares the device's screening count with the oracle rule evaluated on the
DEVICE's own Schwarz bounds (bit-exact; bounds computed on two machines may differ in the last ulp).

Stored per SCF fixture (tests/golden/<name>_scf.npz): E_k (electronic + Vnn) after k = 1..n Roothaan steps from the
core-Hamiltonian guess (HartreeFockModule.jl:73-84,14-41 semantics), all integrals from the oracle.

Run times on 8 cores: benzene seconds, (H2O)32 ~2 min per build, C40H82 ~5 min, (H2O)64/cc-pVDZ ~30 min.
Usage: python tests/golden/make_golden_configs.py [name ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from quantumlab_jl_b200 import workloads  # noqa: E402
from quantumlab_jl_b200.geometry import computeEnergyInteratomicRepulsion  # noqa: E402
from quantumlab_jl_b200.shell import flattenShells  # noqa: E402

JK_CONFIGS = {"benzene_ccpvdz": 1, "h2o32_631gs": 8, "c40h82_def2svp": 16, "h2o64_ccpvdz": 32}
SCF_CONFIGS = {"benzene_ccpvdz": 3, "h2o32_631gs": 3}
TAU = 1e-12


def synthetic_density(shells, seed=7):
    cen = np.array([s.center for s in shells for _ in range((s.lqn.exponent + 1) * (s.lqn.exponent + 2) // 2)])
    d = np.linalg.norm(cen[:, None, :] - cen[None, :, :], axis=2)
    rng = np.random.default_rng(seed)
    A = rng.standard_normal(d.shape) * 0.2 * np.exp(-0.35 * d)
    return A + A.T


def check_vector(n, seed=11):
    return np.random.default_rng(seed).standard_normal(n)


def atoms_of(geo):
    xyz = np.array([a.position for a in geo.atoms], dtype=float)
    Z = np.array([a.element.atomicNumber for a in geo.atoms], dtype=float)
    return xyz, Z


def make_jk(name, stride):
    geo, shells, nocc = workloads.build(name)
    b = orc.Basis(*flattenShells(shells))
    P = synthetic_density(shells)
    t = time.time()
    J, K, counts = orc.md_jk(b, P, tau=TAU)
    cols = np.arange(0, b.N, stride)
    r = check_vector(b.N)
    np.savez_compressed(os.path.join(HERE, name + "_jk.npz"), cols=cols, Jc=J[:, cols], Kc=K[:, cols], Jr=J @ r, Kr=K @ r,
                        sumJ=J.sum(), sumK=K.sum(), trJP=np.trace(J @ P), trKP=np.trace(K @ P), counts=np.array(counts))
    print(f"{name}: N={b.N} kept/unique {counts} trJP {np.trace(J @ P)} trKP {np.trace(K @ P)}  {time.time() - t:.0f} s", flush=True)


def make_scf(name, nsteps):
    geo, shells, nocc = workloads.build(name)
    b = orc.Basis(*flattenShells(shells))
    xyz, Z = atoms_of(geo)
    S = orc.one_electron(b, "overlap")
    T = orc.one_electron(b, "kinetic")
    V = orc.one_electron(b, "nuclear", xyz, Z)
    Vnn = computeEnergyInteratomicRepulsion(geo)
    Q = orc.schwarz(b)
    eps, P = orc.scf_step(T + V, S, nocc)
    E = []
    t = time.time()
    for k in range(nsteps):
        J, K, _ = orc.md_jk(b, P, tau=TAU, Q=Q)
        E.append(orc.energy_hf(P, T, V, J, K) + Vnn)
        print(f"{name}: E after step {k + 1} = {E[-1]:.10f}  ({time.time() - t:.0f} s)", flush=True)
        eps, P = orc.scf_step(T + V + 2 * J - K, S, nocc)
    np.savez_compressed(os.path.join(HERE, name + "_scf.npz"), E=np.array(E), Vnn=Vnn, nocc=nocc)


if __name__ == "__main__":
    names = sys.argv[1:] or list(JK_CONFIGS)
    for n in names:
        if n in JK_CONFIGS:
            make_jk(n, JK_CONFIGS[n])
    for n in names:
        if n in SCF_CONFIGS:
            make_scf(n, SCF_CONFIGS[n])
