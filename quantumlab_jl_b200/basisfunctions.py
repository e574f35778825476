"""PrimitiveGaussianBasisFunction / ContractedGaussianBasisFunction / identityCGF (mirror of BasisFunctionsModule.jl:10-23,
50-51), expandShell (ShellModule.jl:119-128), GaussianBasis / computeBasis / convert (BasisModule.jl:22-24, 51-58, 87-93)
and the regrouping of basis-function lists into device shells.

The device works on complete Cartesian shells.  A list of contracted functions reaches it by regrouping: consecutive
functions that are the Cartesian components of one shell (same centre, same exponents, coefficient vectors that differ
by one factor per function, components in the reference's order) become one shell; any other function becomes a shell
of its own of which only one component is used.  `scale[i]` carries the factor between the i-th function and the
device's component (the device applies the non-axial factors of ShellModule.jl:50-58 itself), so integrals over the
functions are `scale`-weighted picks from the device's arrays -- no integral is ever evaluated on the host.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .base import LQuantumNumber, MQuantumNumber, MQuantumNumbers, Position, doublefactorial, origin
from .shell import Shell, computeFactorsNonAxial


@dataclass
class PrimitiveGaussianBasisFunction:
    """x^a y^b z^c exp(-zeta r^2) at `center` (BasisFunctionsModule.jl:10-14)."""

    center: Position
    exponent: float
    mqn: MQuantumNumber


@dataclass
class ContractedGaussianBasisFunction:
    """sum_i coefficients[i] * primitiveBFs[i] (BasisFunctionsModule.jl:20-23); no extra overall scale."""

    coefficients: list
    primitiveBFs: list


identityPGF = PrimitiveGaussianBasisFunction(origin, 0.0, MQuantumNumber(0, 0, 0))
identityCGF = ContractedGaussianBasisFunction([1.0], [identityPGF])


def expandShell(sh: Shell) -> list[ContractedGaussianBasisFunction]:
    """The nbf(sh) contracted functions of a shell, non-axial factors folded into the coefficients (ShellModule.jl:119-128)."""
    out = []
    for mqn, f in zip(MQuantumNumbers(sh.lqn), computeFactorsNonAxial(sh.lqn)):
        pg = [PrimitiveGaussianBasisFunction(sh.center, float(e), mqn) for e in sh.exponents]
        out.append(ContractedGaussianBasisFunction([float(c) * f for c in sh.coefficients], pg))
    return out


def _self_overlap_prim(zeta_a: float, zeta_b: float, mqn: MQuantumNumber) -> float:
    """<x^a y^b z^c e^{-za r^2} | x^a y^b z^c e^{-zb r^2}> at a common centre (closed form of IntegralsModule.jl:138-163)."""
    g = zeta_a + zeta_b
    L = mqn.x + mqn.y + mqn.z
    df = doublefactorial(2 * mqn.x - 1) * doublefactorial(2 * mqn.y - 1) * doublefactorial(2 * mqn.z - 1)
    return df / (2.0 * g) ** L * (math.pi / g) ** 1.5


@dataclass
class GaussianBasis:
    """A basis as a plain list of contracted functions (BasisModule.jl:22-24)."""

    contractedBFs: list = field(default_factory=list)

    @classmethod
    def from_shells(cls, shells) -> "GaussianBasis":
        """convert(GaussianBasis, shells) (BasisModule.jl:85-93)."""
        if isinstance(shells, Shell):
            shells = [shells]
        bas = cls([])
        for sh in shells:
            bas.contractedBFs.extend(expandShell(sh))
        return bas

    def __len__(self):
        return len(self.contractedBFs)


def computeBasis(basSet, geo, normalize: bool = True) -> GaussianBasis:
    """BasisModule.jl:28-58: geometry atom order -> definition order -> Cartesian component order; each primitive is
    normalised by its self-overlap (renormprimitives!) and then each contracted function by its own (normalize!,
    IntegralsModule.jl:472-475).  Same-centre self-overlaps have a closed form; nothing else is needed here."""
    bas = GaussianBasis([])
    for atom in geo.atoms:
        for cdef in basSet.definitions[atom.element]:
            for mqn in MQuantumNumbers(cdef.lQuantumNumber):
                coeffs = [float(p.prefactor) for p in cdef.primitives]
                pg = [PrimitiveGaussianBasisFunction(atom.position, float(p.exponent), mqn) for p in cdef.primitives]
                if normalize:
                    coeffs = [c / math.sqrt(_self_overlap_prim(p.exponent, p.exponent, mqn)) for c, p in zip(coeffs, pg)]
                    norm = sum(ci * cj * _self_overlap_prim(pi.exponent, pj.exponent, mqn)
                               for ci, pi in zip(coeffs, pg) for cj, pj in zip(coeffs, pg))
                    coeffs = [c / math.sqrt(norm) for c in coeffs]
                bas.contractedBFs.append(ContractedGaussianBasisFunction(coeffs, pg))
    return bas


# ---------------------------------------------------------------------------------------------- regrouping into shells
def _single_centre(cg: ContractedGaussianBasisFunction):
    p0 = cg.primitiveBFs[0]
    for p in cg.primitiveBFs[1:]:
        if tuple(p.center) != tuple(p0.center) or tuple(p.mqn) != tuple(p0.mqn):
            raise ValueError("the device path needs contracted functions whose primitives share centre and Cartesian powers")
    return Position(*p0.center), MQuantumNumber(*p0.mqn)


def shells_from_functions(cgbfs):
    """-> (shells, shell_of[i], comp_of[i], scale[i]): function i == scale[i] * (component comp_of[i] of shells[shell_of[i]])
    as the device defines that component (shell coefficients x non-axial factor)."""
    shells, shell_of, comp_of, scale = [], [], [], []
    i, n = 0, len(cgbfs)
    while i < n:
        center, mqn = _single_centre(cgbfs[i])
        L = mqn.x + mqn.y + mqn.z
        comps = MQuantumNumbers(L)
        nonax = computeFactorsNonAxial(L)
        exps = [p.exponent for p in cgbfs[i].primitiveBFs]
        base = np.array(cgbfs[i].coefficients, dtype=np.float64)
        k0 = comps.index(mqn)
        shells.append(Shell(center, LQuantumNumber(L), exps, base, renorm=False))
        shell_of.append(len(shells) - 1)
        comp_of.append(k0)
        scale.append(1.0 / nonax[k0])
        i += 1
        # absorb the following functions while they are the next components of the same shell
        k = k0 + 1
        while i < n and k < len(comps):
            try:
                c2, m2 = _single_centre(cgbfs[i])
            except ValueError:
                break
            e2 = [p.exponent for p in cgbfs[i].primitiveBFs]
            co2 = np.array(cgbfs[i].coefficients, dtype=np.float64)
            if tuple(c2) != tuple(center) or m2 != comps[k] or e2 != exps or co2.shape != base.shape:
                break
            ratio = co2[0] / base[0] if base[0] != 0.0 else float("nan")
            if not np.allclose(co2, ratio * base, rtol=1e-13, atol=0.0):
                break
            shell_of.append(len(shells) - 1)
            comp_of.append(k)
            scale.append(ratio / nonax[k])
            i += 1
            k += 1
    return shells, shell_of, comp_of, np.array(scale)


class FunctionMap:
    """Index map between a list of contracted functions and the device basis built from `shells_from_functions`."""

    def __init__(self, cgbfs):
        from .shell import nbf

        self.shells, shell_of, comp_of, self.scale = shells_from_functions(list(cgbfs))
        off = np.concatenate([[0], np.cumsum([nbf(s) for s in self.shells])]).astype(int)
        self.index = np.array([off[s] + c for s, c in zip(shell_of, comp_of)], dtype=int)  # device basis-function index
        self.nbf_device = int(off[-1])
        self.n = len(self.index)

    def embed_density(self, P):
        """P over the functions -> the density the device must see."""
        P = np.asarray(P, dtype=np.float64)
        D = np.zeros((self.nbf_device, self.nbf_device), order="F")
        D[np.ix_(self.index, self.index)] = P * np.outer(self.scale, self.scale)
        return D

    def extract_matrix(self, M):
        return M[np.ix_(self.index, self.index)] * np.outer(self.scale, self.scale)

    def extract_tensor(self, T):
        ix = self.index
        s = self.scale
        return T[np.ix_(ix, ix, ix, ix)] * s[:, None, None, None] * s[None, :, None, None] * s[None, None, :, None] * s[None, None, None, :]
