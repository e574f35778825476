"""Shell (mirror of ShellModule.jl) and the basis -> shell list builder (BasisModule.jl:72-83)."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .base import LQuantumNumber, MQuantumNumbers, Position, doublefactorial, numberMQNsCartesian
from .basisset import BasisSet
from .geometry import Geometry


class AbstractShell:
    pass


@dataclass
class Shell(AbstractShell):
    """Shell(center, lqn, exponents, coefficients; renorm=True)  -- ShellModule.jl:25-48.

    With ``renorm`` the coefficients are taken as written in a basis-set file and normalised the
    libint way: primitive-by-primitive (:36) and then the whole contraction (:39-44), so that axial
    functions are unit-normalised.  Unlike the reference the caller's array is not mutated.
    """

    center: Position
    lqn: LQuantumNumber
    exponents: np.ndarray
    coefficients: np.ndarray

    def __init__(self, center, lqn, exponents, coefficients, renorm: bool = True):
        self.center = Position(*map(float, center))
        self.lqn = lqn if isinstance(lqn, LQuantumNumber) else LQuantumNumber(lqn)
        e = np.array(exponents, dtype=np.float64)
        c = np.array(coefficients, dtype=np.float64)
        if e.shape != c.shape or e.ndim != 1 or e.size == 0:
            raise ValueError("exponents and coefficients must be equally long, non-empty vectors")
        L = self.lqn.exponent
        if renorm:
            df = doublefactorial(2 * L - 1)
            c = c * np.sqrt((2.0**L * (2 * e) ** (L + 1.5)) / (math.pi**1.5 * df))
            norm = 0.0
            for p in range(e.size):
                for q in range(e.size):
                    norm += (df * math.pi**1.5 * c[p] * c[q]) / (2.0**L * (e[p] + e[q]) ** (L + 1.5))
            c = c / math.sqrt(norm)
        self.exponents, self.coefficients = e, c


def nbf(shell: Shell) -> int:
    return numberMQNsCartesian(shell.lqn)


def computeFactorsNonAxial(lqn) -> list[float]:
    """ShellModule.jl:50-58: 1/sqrt((2mx-1)!!(2my-1)!!(2mz-1)!!/(2L-1)!!) per Cartesian component."""
    L = lqn.exponent if isinstance(lqn, LQuantumNumber) else int(lqn)
    out = []
    for m in MQuantumNumbers(L):
        na = doublefactorial(2 * m.x - 1) * doublefactorial(2 * m.y - 1) * doublefactorial(2 * m.z - 1) / doublefactorial(2 * L - 1)
        out.append(1.0 / math.sqrt(na))
    return out


def computeBasisShells(basSet: BasisSet, geo: Geometry) -> list[Shell]:
    """Geometry atom order -> basis-set definition order (BasisModule.jl:72-83)."""
    shells = []
    for atom in geo.atoms:
        for cdef in basSet.definitions[atom.element]:
            shells.append(Shell(atom.position, cdef.lQuantumNumber,
                                [p.exponent for p in cdef.primitives], [p.prefactor for p in cdef.primitives]))
    return shells


def flattenShells(shells: list[Shell]):
    """SoA view handed to the C ABI: centers(3*nsh), L, nprim, exps, coefs (already normalised)."""
    nsh = len(shells)
    centers = np.array([c for s in shells for c in s.center], dtype=np.float64).reshape(nsh * 3)
    L = np.array([s.lqn.exponent for s in shells], dtype=np.int32)
    nprim = np.array([s.exponents.size for s in shells], dtype=np.int32)
    exps = np.concatenate([s.exponents for s in shells]).astype(np.float64)
    coefs = np.concatenate([s.coefficients for s in shells]).astype(np.float64)
    return centers, L, nprim, exps, coefs


def totalBasisFunctions(shells: list[Shell]) -> int:
    return sum(nbf(s) for s in shells)
