"""computeMatrixCoulomb / Exchange / Fock and the one-electron matrices on the B200 backend
(mirror of SpecialMatricesModule.jl:102-177, 209-316; argument meaning and order as in the reference).

`basis` may be a `B200Shells` (device handle reused across calls, the intended use inside an SCF loop), a plain
list of `Shell` (a temporary handle is created, as the reference creates and destroys a libint2 engine per call,
SpecialMatricesModule.jl:238-242), a `GaussianBasis` (:210-225, 247-262: regrouped into device shells, see
basisfunctions.py) or, for Coulomb/Exchange, a stored N^4 tensor (:227-229, 264-266; contracted on the device by
`qlb_tensor_contract`, small-N test use).  There is no host arithmetic path.
"""
from __future__ import annotations

import numpy as np

from . import libqlb
from .b200shells import B200Shells
from .basisfunctions import FunctionMap, GaussianBasis


def _as_b200(basis):
    if isinstance(basis, B200Shells):
        return basis, False
    return B200Shells(basis), True


def _tensor_contract(tensor, density, exchange: int) -> np.ndarray:
    n = tensor.shape[0]
    if tensor.shape != (n, n, n, n):
        raise ValueError("ERI tensor must be N x N x N x N")
    T = np.asfortranarray(tensor, dtype=np.float64)
    P = np.asfortranarray(density, dtype=np.float64)
    out = np.empty((n, n), dtype=np.float64, order="F")
    lib = libqlb.init()
    libqlb.check(lib.qlb_tensor_contract(n, libqlb.dptr(T), libqlb.dptr(P), exchange, libqlb.dptr(out)))
    return out


def _function_basis_jk(basis: GaussianBasis, density):
    fmap = FunctionMap(basis.contractedBFs)
    dev = B200Shells(fmap.shells)
    try:
        J, K = dev.build_jk(fmap.embed_density(density))
        return fmap.extract_matrix(J), fmap.extract_matrix(K)
    finally:
        dev.destroy()


def computeMatrixCoulomb(basis, density) -> np.ndarray:
    if isinstance(basis, np.ndarray) and basis.ndim == 4:
        return _tensor_contract(basis, density, 0)
    if isinstance(basis, GaussianBasis):
        return _function_basis_jk(basis, density)[0]
    b, own = _as_b200(basis)
    try:
        return b.jk_cached(density)[0].copy()
    finally:
        if own:
            b.destroy()


def computeMatrixExchange(basis, density) -> np.ndarray:
    if isinstance(basis, np.ndarray) and basis.ndim == 4:
        return _tensor_contract(basis, density, 1)
    if isinstance(basis, GaussianBasis):
        return _function_basis_jk(basis, density)[1]
    b, own = _as_b200(basis)
    try:
        return b.jk_cached(density)[1].copy()
    finally:
        if own:
            b.destroy()


def _one_electron(basis, which, geometry=None) -> np.ndarray:
    if isinstance(basis, GaussianBasis):
        fmap = FunctionMap(basis.contractedBFs)
        dev = B200Shells(fmap.shells)
        try:
            return fmap.extract_matrix(dev.one_electron(which, geometry))
        finally:
            dev.destroy()
    b, own = _as_b200(basis)
    try:
        return b.one_electron(which, geometry)
    finally:
        if own:
            b.destroy()


def computeMatrixOverlap(basis) -> np.ndarray:
    return _one_electron(basis, "overlap")


def computeMatrixKinetic(basis) -> np.ndarray:
    return _one_electron(basis, "kinetic")


def computeMatrixNuclearAttraction(basis, geometry) -> np.ndarray:
    return _one_electron(basis, "nuclear", geometry)


def computeMatrixFock(*args) -> np.ndarray:
    """computeMatrixFock(T, V, J, K)                      -- SpecialMatricesModule.jl:284-290
    computeMatrixFock(basis, geometry, density)         -- :292-303 (T, V, J, K recomputed; J and K in one pass)
    computeMatrixFock(basis, geometry, density, tensorC, tensorX)   -- :305-316 (stored tensors for J and K)"""
    if len(args) == 5:
        basis, geometry, density, tc, tx = args
        H = computeMatrixKinetic(basis) + computeMatrixNuclearAttraction(basis, geometry)
        return H + 2 * computeMatrixCoulomb(tc, density) - computeMatrixExchange(tx, density)
    if len(args) == 4:
        T, V, J, K = args
        return T + V + 2 * J - K
    if len(args) == 3:
        basis, geometry, density = args
        if isinstance(basis, GaussianBasis):
            H = computeMatrixKinetic(basis) + computeMatrixNuclearAttraction(basis, geometry)
            J, K = _function_basis_jk(basis, density)
            return H + 2 * J - K
        b, own = _as_b200(basis)
        try:
            H = b.one_electron("kinetic") + b.one_electron("nuclear", geometry)
            return b.fock(H, density)
        finally:
            if own:
                b.destroy()
    raise TypeError("computeMatrixFock takes (T,V,J,K), (basis,geometry,density) or (basis,geometry,density,tensorC,tensorX)")
