"""Geometry (mirror of GeometryModule.jl:13-84: XYZ reader, the Angstrom->bohr constant of :55,
interatomic repulsion :64-80, electron count :82-84)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field

from .atom import Atom, Element
from .base import Position, distance

BOHR_IN_ANGSTROM = 0.529177210  # GeometryModule.jl:55


def angstrom2bohr(x):
    if isinstance(x, Position):
        return Position(*(c / BOHR_IN_ANGSTROM for c in x))
    return x / BOHR_IN_ANGSTROM


def bohr2angstrom(x: float) -> float:
    return x * BOHR_IN_ANGSTROM


@dataclass
class Geometry:
    atoms: list[Atom] = field(default_factory=list)

    @classmethod
    def from_file(cls, filename: str) -> "Geometry":
        for cand in (filename, filename + ".xyz", filename + ".XYZ"):
            if os.path.isfile(cand):
                return readGeometryXYZ(cand)
        raise FileNotFoundError(f"Couldn't find file: {filename}")


def readGeometryXYZ(filename: str, unit: str = "Bohr") -> Geometry:
    with open(filename) as fh:
        lines = fh.read().splitlines()
    return parseGeometryXYZ(lines, unit)


def parseGeometryXYZ(lines: list[str], unit: str = "Bohr") -> Geometry:
    geo = Geometry([])
    for line in lines[2:]:  # line 1 = atom count, line 2 = comment
        cols = line.split()
        if len(cols) < 4:
            continue
        pos = Position(float(cols[1]), float(cols[2]), float(cols[3]))
        if unit == "Bohr":
            pos = angstrom2bohr(pos)
        geo.atoms.append(Atom(Element(cols[0]), pos))
    return geo


def geometryFromAngstrom(symbols, coords_angstrom) -> Geometry:
    return Geometry([Atom(Element(s), angstrom2bohr(Position(*map(float, c)))) for s, c in zip(symbols, coords_angstrom)])


def computeEnergyInteratomicRepulsion(geo: Geometry) -> float:
    result = 0.0
    for a in range(len(geo.atoms)):
        for b in range(a):
            A, B = geo.atoms[a], geo.atoms[b]
            result += A.element.atomicNumber * B.element.atomicNumber / distance(A.position, B.position)
    return result


def computeNumberElectrons(geo: Geometry, charge: int = 0) -> int:
    return sum(at.element.atomicNumber for at in geo.atoms) - charge
