"""Initial density guesses (mirror of InitialGuessModule.jl:5-16, 95-105).

ZeroGuess / SADGuess singletons and the block-diagonal superposition-of-atomic-densities builder.  Only
the STO-3G H/O tables (InitialGuessModule.jl:31-36, 61-66) are carried: they are the data behind the
reference's Fock known-answer test (test/runtests.jl:148) and the test-suite's SCF start density.
"""
from __future__ import annotations

import numpy as np


class InitialGuess:
    pass


class _ZeroGuess(InitialGuess):
    def __repr__(self):
        return "ZeroGuess"


class _SADGuess(InitialGuess):
    def __repr__(self):
        return "SADGuess"


ZeroGuess = _ZeroGuess()
SADGuess = _SADGuess()

_O_A = np.zeros((5, 5))
_O_A[0, 0] = _O_A[1, 1] = 1.065739
_O_A[0, 1] = _O_A[1, 0] = -0.264689
_O_A[2, 2] = _O_A[3, 3] = 1.0
_O_B = np.zeros((5, 5))
_O_B[:2, :2] = _O_A[:2, :2]
SAD_STO3G_ALPHA = {"H": np.array([[1.0]]), "O": _O_A}
SAD_STO3G_BETA = {"H": np.array([[0.0]]), "O": _O_B}
_TABLES = {"STO-3G": (SAD_STO3G_ALPHA, SAD_STO3G_BETA)}


def sad_density(geometry, table) -> np.ndarray:
    blocks = [table[a.element.symbol] for a in geometry.atoms]
    n = sum(b.shape[0] for b in blocks)
    out = np.zeros((n, n))
    o = 0
    for b in blocks:
        out[o:o + b.shape[0], o:o + b.shape[0]] = b
        o += b.shape[0]
    return out


def computeDensityGuessSAD(methodName: str, basisSetName: str, geometry):
    if methodName not in ("HF", "HartreeFock"):
        raise ValueError("No SAD guess implemented for method " + methodName)
    a, b = _TABLES[basisSetName.upper()]
    return sad_density(geometry, a), sad_density(geometry, b)


def computeDensityMatrixGuess(guess, nbf: int, basisSetName: str | None = None, geometry=None) -> np.ndarray:
    """InitialGuessModule.jl:11-16: ZeroGuess -> zeros, SADGuess -> (alpha+beta)/2, a matrix -> itself."""
    if guess is ZeroGuess:
        return np.zeros((nbf, nbf))
    if guess is SADGuess:
        a, b = computeDensityGuessSAD("HF", basisSetName, geometry)
        return (a + b) / 2
    return np.asarray(guess, dtype=np.float64)
