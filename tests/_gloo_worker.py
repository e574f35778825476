"""Worker for test_multirank_gloo.py: world-size-2 rehearsal (gloo, CPU) of the multi-GPU host logic.

Each rank evaluates ITS share of the unique quartets (the same round-robin deal of bra pairs the device path uses,
computed here by the CPU oracle standing in for the kernels), the partial [J;K] buffers are summed with ONE all-reduce,
and every rank checks the result against the full build."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from quantumlab_jl_b200 import workloads  # noqa: E402
from quantumlab_jl_b200.shell import flattenShells  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    geo, shells, nocc = workloads.build("h2o2_631gs")
    b = orc.Basis(*flattenShells(shells))
    rng = np.random.default_rng(4)
    A = rng.standard_normal((b.N, b.N)) * 0.2
    P = A + A.T
    Q = orc.schwarz(b)
    Jp, Kp, cp = orc.md_jk(b, P, tau=1e-12, Q=Q, rank=rank, nranks=world)
    buf = torch.from_numpy(np.concatenate([Jp.reshape(-1, order="F"), Kp.reshape(-1, order="F")]))
    counts = torch.tensor([cp[0], cp[1]], dtype=torch.int64)
    dist.all_reduce(buf)      # the single collective of a build: sum of the partial [J;K]
    dist.all_reduce(counts)
    J = buf[: b.N * b.N].numpy().reshape((b.N, b.N), order="F")
    K = buf[b.N * b.N:].numpy().reshape((b.N, b.N), order="F")
    Jf, Kf, cf = orc.md_jk(b, P, tau=1e-12, Q=Q)
    assert np.abs(J - Jf).max() < 1e-12 and np.abs(K - Kf).max() < 1e-12
    assert (int(counts[0]), int(counts[1])) == cf, (counts, cf)
    assert 0 < cp[0] < cf[0]  # the rank really held only a share
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
