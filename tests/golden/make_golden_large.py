"""Golden sample of J/K for the headline workload (H2O)_32 / 6-31G* (N = 608), from the CPU oracle's MD path
(8-fold symmetry, same integer screening, tau = 1e-12).  Takes ~1-2 minutes on 8 cores.  The density is a seeded,
spatially decaying symmetric matrix (no SCF needed).  Stored: every 37th column of J and K plus full-matrix checksums.
Usage: python tests/golden/make_golden_large.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from quantumlab_jl_b200 import workloads  # noqa: E402
from quantumlab_jl_b200.shell import flattenShells  # noqa: E402


def synthetic_density(shells, seed=7):
    cen = np.array([s.center for s in shells for _ in range((s.lqn.exponent + 1) * (s.lqn.exponent + 2) // 2)])
    d = np.linalg.norm(cen[:, None, :] - cen[None, :, :], axis=2)
    rng = np.random.default_rng(seed)
    A = rng.standard_normal(d.shape) * 0.2 * np.exp(-0.35 * d)
    return A + A.T


def main():
    geo, shells, nocc = workloads.build("h2o32_631gs")
    b = orc.Basis(*flattenShells(shells))
    P = synthetic_density(shells)
    J, K, counts = orc.md_jk(b, P, tau=1e-12)
    cols = np.arange(0, b.N, 37)
    np.savez_compressed(os.path.join(HERE, "h2o32_631gs_jk_sample.npz"), cols=cols, Jc=J[:, cols], Kc=K[:, cols],
                        sumJ=J.sum(), sumK=K.sum(), trJP=np.trace(J @ P), trKP=np.trace(K @ P), counts=np.array(counts))
    print("kept/unique", counts, "trJP", np.trace(J @ P), "trKP", np.trace(K @ P))


if __name__ == "__main__":
    main()
