"""Host-side input layer (mirror of BasisSetModule / GeometryModule / BaseModule): grammar, ordering, constants."""
import math

import numpy as np
import pytest

from quantumlab_jl_b200.base import LQuantumNumber, MQuantumNumbers, doublefactorial, numberMQNsCartesian
from quantumlab_jl_b200.basisset import BasisSetTX93, loadBasisSet
from quantumlab_jl_b200.geometry import (BOHR_IN_ANGSTROM, computeEnergyInteratomicRepulsion, computeNumberElectrons,
                                         parseGeometryXYZ)
from quantumlab_jl_b200.shell import Shell, computeBasisShells, computeFactorsNonAxial

TX93 = """! STO-3G
FOR O
S    130.7093200   0.15432897
      23.8088610   0.53532814
       6.4436083   0.44463454
SP     5.0331513  -0.09996723   0.15591627
       1.1695961   0.39951283   0.60768372
       0.3803890   0.70011547   0.39195739
FOR H
S      3.42525091  0.15432897
       0.62391373  0.53532814
       0.16885540  0.44463454
"""


def test_tx93_grammar_sp_split_and_order():
    # BasisSetModule.jl:54-98: SP rows carry two coefficient columns and become an S then a P contraction
    bs = BasisSetTX93(TX93)
    o = [d for e, d in bs.definitions.items() if e.symbol == "O"][0]
    assert [c.lQuantumNumber.symbol for c in o] == ["S", "S", "P"]
    assert [len(c.primitives) for c in o] == [3, 3, 3]
    assert o[1].primitives[0].prefactor == -0.09996723 and o[2].primitives[0].prefactor == 0.15591627
    assert o[1].primitives[2].exponent == 0.3803890 == o[2].primitives[2].exponent


def test_tx93_errors():
    with pytest.raises(ValueError):
        BasisSetTX93("S 1.0 1.0\n")  # contraction before any FOR header
    with pytest.raises(ValueError):
        BasisSetTX93("FOR H\nSP 1.0 1.0\n")  # SP needs two coefficient columns
    with pytest.raises(FileNotFoundError):
        loadBasisSet("no-such-basis")


def test_vendored_basis_sets_shell_structure():
    # shells per atom as SURVEY.md 6 derives them (s/p/d counts; SP split)
    expect = {"sto-3g": {"H": "S", "O": "SSP", "C": "SSP"}, "6-31g*": {"H": "SS", "O": "SSPSPD"},
              "cc-pvdz": {"H": "SSP", "O": "SSSPPD", "C": "SSSPPD"}, "def2-svp": {"H": "SSP", "C": "SSSPPD"}}
    for name, per_el in expect.items():
        bs = loadBasisSet(name)
        for e, defs in bs.definitions.items():
            if e.symbol in per_el:
                assert "".join(c.lQuantumNumber.symbol for c in defs) == per_el[e.symbol], (name, e.symbol)


def test_cartesian_order_and_counts():
    # BaseModule.jl:110-122: x descending, then y descending
    assert [tuple(m) for m in MQuantumNumbers(1)] == [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    assert [tuple(m) for m in MQuantumNumbers(LQuantumNumber("D"))] == [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]
    assert [numberMQNsCartesian(l) for l in range(5)] == [1, 3, 6, 10, 15]
    assert doublefactorial(5) == 15 and doublefactorial(-1) == 1 and doublefactorial(0) == 1
    f = computeFactorsNonAxial(2)  # ShellModule.jl:50-58: d_xy, d_xz, d_yz carry sqrt(3)
    assert np.allclose(f, [1, math.sqrt(3), math.sqrt(3), 1, math.sqrt(3), 1])


def test_xyz_reader_and_repulsion():
    lines = ["3", "water", "H 0.751 0.194 0.0", "O 0.0 -0.388 0.0", "H -0.751 0.194 0.0"]
    geo = parseGeometryXYZ(lines)  # Angstrom -> bohr with 0.529177210 (GeometryModule.jl:55)
    assert BOHR_IN_ANGSTROM == 0.529177210
    assert abs(geo.atoms[0].position.x - 0.751 / 0.529177210) < 1e-15
    assert computeNumberElectrons(geo) == 10 and computeNumberElectrons(geo, 1) == 9
    assert abs(computeEnergyInteratomicRepulsion(geo) - 9.2636625634) < 1e-8  # BASELINE.md derived golden


def test_shell_constructor_contract():
    from quantumlab_jl_b200.base import origin

    raw = np.array([0.1, 0.2, 0.3])
    sh = Shell(origin, LQuantumNumber("S"), [1.0, 2.0, 3.0], raw)
    assert abs(sh.coefficients[1] - 0.41030724) < 1e-8          # runtests.jl:93
    assert np.array_equal(raw, [0.1, 0.2, 0.3])                 # unlike ShellModule.jl:36 the caller's array is untouched
    keep = Shell(origin, LQuantumNumber("P"), [1.0], [0.7], renorm=False)
    assert keep.coefficients[0] == 0.7
    with pytest.raises(ValueError):
        Shell(origin, LQuantumNumber("S"), [1.0, 2.0], [1.0])
    geo = parseGeometryXYZ(["1", "", "O 0 0 0"])
    shells = computeBasisShells(BasisSetTX93(TX93), geo)
    assert [s.lqn.symbol for s in shells] == ["S", "S", "P"]


def test_function_lists_regroup_into_device_shells():
    """a3 / a6: basis-function lists (expandShell, computeBasis) regroup into exactly the shells of computeBasisShells; any
    other list becomes one shell per function with the right component and scale (basisfunctions.py)."""
    import numpy as np

    from quantumlab_jl_b200 import workloads
    from quantumlab_jl_b200.basisfunctions import FunctionMap, GaussianBasis, computeBasis, expandShell, identityCGF
    from quantumlab_jl_b200.basisset import loadBasisSet
    from quantumlab_jl_b200.shell import nbf

    geo, shells, nocc = workloads.build("h2o_631gs")
    bas = GaussianBasis.from_shells(shells)
    assert len(bas) == sum(nbf(s) for s in shells) == 19
    fm = FunctionMap(bas.contractedBFs)
    assert len(fm.shells) == len(shells) and np.allclose(fm.scale, 1.0) and list(fm.index) == list(range(19))
    dsh = [s for s in shells if s.lqn.exponent == 2][0]
    d = expandShell(dsh)  # the oxygen d shell: xy, xz, yz carry sqrt(3)
    assert np.isclose(d[1].coefficients[0] / d[0].coefficients[0], np.sqrt(3.0))
    bas2 = computeBasis(loadBasisSet("6-31g*"), geo)
    fm2 = FunctionMap(bas2.contractedBFs)
    assert len(fm2.shells) == len(shells) and np.allclose(fm2.scale, 1.0, rtol=1e-12)
    # an arbitrary pick of functions: one device shell each, the picked component, scale 1/nonaxial
    pick = [bas.contractedBFs[i] for i in (0, 5, 16, 14)]
    fm3 = FunctionMap(pick)
    assert len(fm3.shells) == 4
    assert identityCGF.primitiveBFs[0].exponent == 0.0 and identityCGF.coefficients == [1.0]


def test_orbital_f_shell_is_rejected_by_name():
    """QLB_MAX_L = 2 on the orbital index: the host mirror names the offending shell before any device call."""
    import pytest

    from quantumlab_jl_b200.b200shells import B200Shells
    from quantumlab_jl_b200.base import LQuantumNumber, Position
    from quantumlab_jl_b200.shell import Shell

    s = Shell(Position(0.0, 0.0, 0.0), LQuantumNumber("S"), [1.0], [1.0])
    f = Shell(Position(0.0, 0.0, 1.0), LQuantumNumber("F"), [0.8], [1.0])
    with pytest.raises(ValueError, match="shell 1 .L = 3"):
        B200Shells([s, f])
