"""Element / Atom (mirror of the parts of AtomModule.jl the hot path touches: symbol -> Z)."""
from __future__ import annotations

from dataclasses import dataclass

from .base import Position

_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
            "Ga Ge As Se Br Kr").split()
_Z = {s.upper(): i + 1 for i, s in enumerate(_SYMBOLS)}


@dataclass(frozen=True)
class Element:
    symbol: str

    def __post_init__(self):
        s = self.symbol.strip()
        if s.upper() not in _Z:
            raise ValueError(f"unknown element {self.symbol!r}")
        object.__setattr__(self, "symbol", s[0].upper() + s[1:].lower())

    @property
    def atomicNumber(self) -> int:
        return _Z[self.symbol.upper()]


@dataclass
class Atom:
    element: Element
    position: Position
