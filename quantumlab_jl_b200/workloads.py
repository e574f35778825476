"""Deterministic synthetic inputs for the BASELINE.json configurations."""
from __future__ import annotations

import math

from .geometry import Geometry, geometryFromAngstrom

# test/h2o.xyz of the reference (Angstrom)
_MONOMER = (("H", (0.751, 0.194, 0.0)), ("O", (0.0, -0.388, 0.0)), ("H", (-0.751, 0.194, 0.0)))


def water() -> Geometry:
    return geometryFromAngstrom([s for s, _ in _MONOMER], [c for _, c in _MONOMER])


def water_cluster(nx: int, ny: int, nz: int, spacing: float = 3.0) -> Geometry:
    """Monomers translated onto a simple-cubic grid, x fastest, then y, then z; no rotation."""
    sym, xyz = [], []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                for s, c in _MONOMER:
                    sym.append(s)
                    xyz.append((c[0] + i * spacing, c[1] + j * spacing, c[2] + k * spacing))
    return geometryFromAngstrom(sym, xyz)


def benzene() -> Geometry:
    """D6h in the xy plane, r_CC = 1.39 A, r_CH = 1.09 A, atom order C1..C6, H1..H6."""
    sym, xyz = [], []
    for r, s in ((1.39, "C"), (1.39 + 1.09, "H")):
        for k in range(6):
            a = math.pi / 3 * k
            sym.append(s)
            xyz.append((r * math.cos(a), r * math.sin(a), 0.0))
    return geometryFromAngstrom(sym, xyz)


def alkane(n: int) -> Geometry:
    """All-trans C_nH_(2n+2): r_CC 1.54 A, r_CH 1.09 A, tetrahedral angles; carbons first, then hydrogens."""
    th = math.radians(109.4712)
    dx, dz = 1.54 * math.sin(th / 2), 1.54 * math.cos(th / 2)
    cs = [(i * dx, 0.0, (i % 2) * dz) for i in range(n)]
    hs = []
    hy, hz = 1.09 * math.sin(th / 2), 1.09 * math.cos(th / 2)
    for i, (x, y, z) in enumerate(cs):
        sgn = -1.0 if i % 2 == 0 else 1.0
        hs.append((x, hy, z + sgn * hz))
        hs.append((x, -hy, z + sgn * hz))
    # chain ends: one more H each, continuing the zig-zag
    x0, _, z0 = cs[0]
    hs.append((x0 - 1.09 * math.sin(th / 2), 0.0, z0 + 1.09 * math.cos(th / 2)))
    xn, _, zn = cs[-1]
    sgn = 1.0 if (n - 1) % 2 == 0 else -1.0
    hs.append((xn + 1.09 * math.sin(th / 2), 0.0, zn + sgn * 1.09 * math.cos(th / 2)))
    return geometryFromAngstrom(["C"] * n + ["H"] * len(hs), cs + hs)


CONFIGS = {
    "h2o_sto3g": dict(geometry=water, basis="sto-3g", nocc=5),
    "benzene_ccpvdz": dict(geometry=benzene, basis="cc-pvdz", nocc=21),
    "h2o32_631gs": dict(geometry=lambda: water_cluster(4, 4, 2), basis="6-31g*", nocc=160),
    "c40h82_def2svp": dict(geometry=lambda: alkane(40), basis="def2-svp", nocc=161),
    "h2o64_ccpvdz": dict(geometry=lambda: water_cluster(4, 4, 4), basis="cc-pvdz", nocc=320),
    # small members of the same families, used by parity tests
    "h2o_631gs": dict(geometry=water, basis="6-31g*", nocc=5),
    "h2o2_631gs": dict(geometry=lambda: water_cluster(2, 1, 1), basis="6-31g*", nocc=10),
    "h2o4_631gs": dict(geometry=lambda: water_cluster(2, 2, 1), basis="6-31g*", nocc=20),
    "h2o_ccpvdz": dict(geometry=water, basis="cc-pvdz", nocc=5),
    "c2h6_def2svp": dict(geometry=lambda: alkane(2), basis="def2-svp", nocc=9),
}


def build(name: str):
    """-> (geometry, shells, nocc) for a named configuration."""
    from .basisset import loadBasisSet
    from .shell import computeBasisShells

    cfg = CONFIGS[name]
    geo = cfg["geometry"]()
    shells = computeBasisShells(loadBasisSet(cfg["basis"]), geo)
    return geo, shells, cfg["nocc"]


def jkfit_aux(geo: Geometry, name: str = "jkfit-et-dz"):
    """Auxiliary shells of config 5: the vendored JK-fitting set (an even-tempered stand-in for cc-pVDZ-JKFIT with s, p, d
    and f shells, data/basis/jkfit-et-dz.tx93) on the atoms of `geo`."""
    from .basisset import loadBasisSet
    from .shell import computeBasisShells

    return computeBasisShells(loadBasisSet(name), geo)


def even_tempered_aux(geo: Geometry):
    """A synthetic auxiliary basis (the cc-pVDZ-JKFIT data cannot be vendored offline): per atom even-tempered
    uncontracted s (6), p (4) and d (3) shells -> 36 Cartesian functions per atom."""
    from .base import LQuantumNumber
    from .shell import Shell

    shells = []
    for atom in geo.atoms:
        for lsym, a0, n in (("S", 0.10, 6), ("P", 0.20, 4), ("D", 0.30, 3)):
            for k in range(n):
                shells.append(Shell(atom.position, LQuantumNumber(lsym), [a0 * 3.0 ** k], [1.0]))
    return shells
