"""Generates the golden vectors under tests/golden/ from the CPU oracle (run here, committed; the GPU box has no
/root/reference and needs none).  The oracle itself is pinned against the reference's known-answer tests by
tests/test_oracle_kat.py.  Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from quantumlab_jl_b200 import workloads  # noqa: E402
from quantumlab_jl_b200.shell import flattenShells  # noqa: E402


def basis(name):
    geo, shells, nocc = workloads.build(name)
    return geo, shells, orc.Basis(*flattenShells(shells)), nocc


def main():
    # 1. H2O/STO-3G: the full ERI tensor by the literal reference formula (Cook closed form, exact Boys)
    geo, shells, b, nocc = basis("h2o_sto3g")
    np.savez_compressed(os.path.join(HERE, "h2o_sto3g_eri.npz"), eri=orc.eri_tensor(b, literal=True))
    # 2. H2O/6-31G* (s,p,d): J, K for a seeded density; every L-class occurs; Cook closed form via the N^4 tensor
    geo, shells, b, nocc = basis("h2o_631gs")
    rng = np.random.default_rng(2026)
    A = rng.standard_normal((b.N, b.N)) * 0.25
    P = A + A.T
    eri = orc.eri_tensor(b, literal=False)
    J = orc.coulomb_from_tensor(eri, P)
    K = orc.exchange_from_tensor(eri, P)
    # selected d-containing blocks, indices (shell i,j,k,l)
    blocks = {}
    dsh = [i for i, s in enumerate(shells) if s.lqn.exponent == 2]
    psh = [i for i, s in enumerate(shells) if s.lqn.exponent == 1]
    ssh = [i for i, s in enumerate(shells) if s.lqn.exponent == 0]
    for q in [(dsh[0], dsh[0], dsh[0], dsh[0]), (dsh[0], psh[1], dsh[0], psh[0]), (dsh[0], dsh[0], psh[0], ssh[5]),
              (psh[0], dsh[0], ssh[0], psh[1]), (ssh[6], psh[0], dsh[0], ssh[1])]:
        blocks["blk_%d_%d_%d_%d" % q] = orc.eri_block(b, *q, literal=False)
    np.savez_compressed(os.path.join(HERE, "h2o_631gs_jk.npz"), P=P, J=J, K=K, **blocks)
    print("golden vectors written")


if __name__ == "__main__":
    main()
