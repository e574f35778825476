"""ERI entry points of IntegralsModule / SpecialMatricesModule on the B200 backend.

computeElectronRepulsionIntegral(mu, nu, la, si)               -- IntegralsModule.jl:412-445 (four primitives) and
    :447-462 (four contracted functions): (mu nu|la si), Mulliken order, no normalisation beyond the coefficients.
computeTensorBlockElectronRepulsionIntegrals(...)              -- four Shells (SpecialMatricesModule.jl:184-186 /
    LibInt2Module.jl:319-334) or four vectors of contracted functions (IntegralsModule.jl:464-470): dense block.
computeTensorElectronRepulsionIntegrals(basis)                 -- SpecialMatricesModule.jl:180-182 (GaussianBasis) and
    :198-207 (shells): the N^4 tensor (small N).
Every integral is evaluated by the class kernels behind `qlb_eri_tensor` / `qlb_eri_block`; the host only regroups
functions into shells and picks / rescales components (basisfunctions.py).
"""
from __future__ import annotations

import numpy as np

from .b200shells import B200Shells
from .basisfunctions import (ContractedGaussianBasisFunction, FunctionMap, GaussianBasis, PrimitiveGaussianBasisFunction)
from .shell import Shell


def _as_cgbf(f) -> ContractedGaussianBasisFunction:
    if isinstance(f, PrimitiveGaussianBasisFunction):
        return ContractedGaussianBasisFunction([1.0], [f])
    if isinstance(f, ContractedGaussianBasisFunction):
        return f
    raise TypeError(f"expected a Primitive- or ContractedGaussianBasisFunction, got {type(f).__name__}")


def _function_tensor(cgbfs) -> np.ndarray:
    fmap = FunctionMap(cgbfs)
    dev = B200Shells(fmap.shells)
    try:
        return fmap.extract_tensor(dev.eri_tensor())
    finally:
        dev.destroy()


def computeElectronRepulsionIntegral(mu, nu, la, si) -> float:
    """(mu nu|la si) over four primitive or four contracted Gaussian functions."""
    T = _function_tensor([_as_cgbf(f) for f in (mu, nu, la, si)])
    return float(T[0, 1, 2, 3])


def computeTensorBlockElectronRepulsionIntegrals(sh1, sh2, sh3, sh4) -> np.ndarray:
    if all(isinstance(s, Shell) for s in (sh1, sh2, sh3, sh4)):
        tmp = B200Shells([sh1, sh2, sh3, sh4])
        try:
            return tmp.eri_block(0, 1, 2, 3)
        finally:
            tmp.destroy()
    # four vectors of contracted functions (IntegralsModule.jl:464-470)
    groups = [[_as_cgbf(f) for f in g] for g in (sh1, sh2, sh3, sh4)]
    flat = [f for g in groups for f in g]
    T = _function_tensor(flat)
    off = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(int)
    return T[off[0]:off[1], off[1]:off[2], off[2]:off[3], off[3]:off[4]].copy(order="F")


def computeTensorElectronRepulsionIntegrals(basis) -> np.ndarray:
    if isinstance(basis, GaussianBasis):
        return np.asfortranarray(_function_tensor(basis.contractedBFs))
    own = not isinstance(basis, B200Shells)
    b = B200Shells(basis) if own else basis
    try:
        return b.eri_tensor()
    finally:
        if own:
            b.destroy()
