"""computeMatrixCoulomb / Exchange / Fock and the one-electron matrices on the B200 backend
(mirror of SpecialMatricesModule.jl:102-177, 209-316; argument meaning and order as in the reference).

`basis` may be a `B200Shells` (device handle reused across calls, the intended use inside an SCF loop), a plain
list of `Shell` (a temporary handle is created, as the reference creates and destroys a libint2 engine per call,
SpecialMatricesModule.jl:238-242) or, for Coulomb/Exchange, a stored N^4 tensor (SpecialMatricesModule.jl:227-229,
264-266; small-N test use).
"""
from __future__ import annotations

import numpy as np

from .b200shells import B200Shells


def _as_b200(basis):
    if isinstance(basis, B200Shells):
        return basis, False
    return B200Shells(basis), True


def computeMatrixCoulomb(basis, density) -> np.ndarray:
    if isinstance(basis, np.ndarray) and basis.ndim == 4:
        return np.einsum("mnls,ls->mn", basis, density)
    b, own = _as_b200(basis)
    try:
        return b.jk_cached(density)[0].copy()
    finally:
        if own:
            b.destroy()


def computeMatrixExchange(basis, density) -> np.ndarray:
    if isinstance(basis, np.ndarray) and basis.ndim == 4:
        return np.einsum("mlns,ls->mn", basis, density)
    b, own = _as_b200(basis)
    try:
        return b.jk_cached(density)[1].copy()
    finally:
        if own:
            b.destroy()


def computeMatrixOverlap(basis) -> np.ndarray:
    b, own = _as_b200(basis)
    try:
        return b.one_electron("overlap")
    finally:
        if own:
            b.destroy()


def computeMatrixKinetic(basis) -> np.ndarray:
    b, own = _as_b200(basis)
    try:
        return b.one_electron("kinetic")
    finally:
        if own:
            b.destroy()


def computeMatrixNuclearAttraction(basis, geometry) -> np.ndarray:
    b, own = _as_b200(basis)
    try:
        return b.one_electron("nuclear", geometry)
    finally:
        if own:
            b.destroy()


def computeMatrixFock(*args) -> np.ndarray:
    """computeMatrixFock(T, V, J, K)                      -- SpecialMatricesModule.jl:284-290
    computeMatrixFock(basis, geometry, density)         -- :292-303 (T, V, J, K recomputed; J and K in one pass)"""
    if len(args) == 4:
        T, V, J, K = args
        return T + V + 2 * J - K
    if len(args) == 3:
        basis, geometry, density = args
        b, own = _as_b200(basis)
        try:
            H = b.one_electron("kinetic") + b.one_electron("nuclear", geometry)
            return b.fock(H, density)
        finally:
            if own:
                b.destroy()
    raise TypeError("computeMatrixFock takes (T,V,J,K) or (basis,geometry,density)")
