"""BasisSet + TX93/PQS reader (mirror of BasisSetModule.jl:54-117).

Grammar honoured: ``!`` comment lines, ``FOR <El>`` headers, contraction blocks introduced by
``S|P|D|F|G|H|SP|L|SPD`` followed by rows ``exponent coeff[ coeff...]``.  ``SP``/``L`` rows carry two
coefficient columns and are split into an S then a P shell, ``SPD`` into S, P, D
(BasisSetModule.jl:76-98).  Basis-set exchange is unreachable offline, so ``BasisSet(name)`` resolves
to the files vendored under ``data/basis`` (BasisSetExchangeModule.jl:73-95 falls back to a download).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field

from .atom import Element
from .base import LQuantumNumber

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "basis")
_LQN_TOKENS = ("SPD", "SP", "L", "S", "P", "D", "F", "G", "H")


@dataclass
class PrimitiveGaussianBasisFunctionDefinition:
    exponent: float
    prefactor: float


@dataclass
class ContractedGaussianBasisFunctionDefinition:
    lQuantumNumber: LQuantumNumber
    primitives: list[PrimitiveGaussianBasisFunctionDefinition]


@dataclass
class BasisSet:
    definitions: dict[Element, list[ContractedGaussianBasisFunctionDefinition]] = field(default_factory=dict)
    name: str | None = None  # set by loadBasisSet: lets evaluateSCF find the SAD tables (InitialGuessModule.jl:95-105)

    def __class_getitem__(cls, name):  # BasisSet["sto-3g"]
        return loadBasisSet(name)


def _float(tok: str) -> float:
    return float(tok.replace("D", "E").replace("d", "e"))


def BasisSetTX93(text: str) -> BasisSet:
    defs: dict[Element, list[ContractedGaussianBasisFunctionDefinition]] = {}
    element = None
    cur_lqn, cur_rows = None, []

    def flush():
        nonlocal cur_lqn, cur_rows
        if cur_lqn is None:
            return
        if not cur_rows:
            raise ValueError(f"contraction {cur_lqn} without primitives")
        ncoef = len(cur_rows[0]) - 1
        split = {"SPD": ("S", "P", "D"), "SP": ("S", "P"), "L": ("S", "P")}.get(cur_lqn, (cur_lqn,))
        if ncoef != len(split):
            raise ValueError(f"{cur_lqn} block needs {len(split)} coefficient column(s), found {ncoef}")
        for col, sym in enumerate(split):
            prims = [PrimitiveGaussianBasisFunctionDefinition(r[0], r[1 + col]) for r in cur_rows]
            defs[element].append(ContractedGaussianBasisFunctionDefinition(LQuantumNumber(sym), prims))
        cur_lqn, cur_rows = None, []

    for raw in text.splitlines():
        line = raw.rstrip()
        if not line.strip() or line.lstrip().startswith("!"):
            continue
        m = re.match(r"^FOR\s+([A-Za-z]+)\s*$", line.strip())
        if m:
            flush()
            element = Element(m.group(1))
            defs[element] = []
            continue
        toks = line.split()
        if toks[0].upper() in _LQN_TOKENS:
            flush()
            if element is None:
                raise ValueError("contraction before any FOR header")
            cur_lqn = toks[0].upper()
            toks = toks[1:]
        if cur_lqn is None:
            raise ValueError(f"unexpected line {raw!r}")
        if toks:
            cur_rows.append([_float(t) for t in toks])
    flush()
    return BasisSet(defs)


def readBasisSetTX93(filename: str) -> BasisSet:
    with open(filename) as fh:
        return BasisSetTX93(fh.read())


_ALIASES = {"6-31g*": "6-31g_st", "6-31g(d)": "6-31g_st", "sto3g": "sto-3g", "def2svp": "def2-svp", "ccpvdz": "cc-pvdz"}


def loadBasisSet(name: str) -> BasisSet:
    key = name.lower()
    key = _ALIASES.get(key, key)
    for cand in (name, os.path.join(_DATA, key + ".tx93")):
        if os.path.isfile(cand):
            bs = readBasisSetTX93(cand)
            bs.name = {"6-31g_st": "6-31G*"}.get(key, key.upper())
            return bs
    raise FileNotFoundError(f"basis set {name!r} is not vendored under {_DATA} (no network: see SURVEY.md 7.1)")
