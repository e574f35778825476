"""The synthetic BASELINE.json workloads have the sizes SURVEY.md 6 derives (Cartesian functions, SP shells split)."""
import pytest

from quantumlab_jl_b200 import workloads
from quantumlab_jl_b200.shell import totalBasisFunctions


@pytest.mark.parametrize("name,natoms,nsh,nbf,nocc", [
    ("h2o_sto3g", 3, 5, 7, 5),
    ("benzene_ccpvdz", 12, 54, 120, 21),
    ("h2o32_631gs", 96, 320, 608, 160),
    ("c40h82_def2svp", 122, 486, 1010, 161),
    ("h2o64_ccpvdz", 192, 768, 1600, 320),
])
def test_workload_sizes(name, natoms, nsh, nbf, nocc):
    geo, shells, n = workloads.build(name)
    assert len(geo.atoms) == natoms and len(shells) == nsh and totalBasisFunctions(shells) == nbf and n == nocc
    assert sum(a.element.atomicNumber for a in geo.atoms) == 2 * nocc  # closed shell


def test_shell_order_follows_atoms_then_definition():
    geo, shells, _ = workloads.build("h2o_sto3g")
    # H1s, O1s, O2s, O2p, H1s  (BasisModule.jl:72-83; SP split into S then P, BasisSetModule.jl:80-90)
    assert [s.lqn.symbol for s in shells] == ["S", "S", "S", "P", "S"]
    assert shells[0].center == geo.atoms[0].position and shells[3].center == geo.atoms[1].position


def test_water_geometry_matches_reference_file():
    geo = workloads.water()
    # test/h2o.xyz converted with 0.529177210 (GeometryModule.jl:55)
    assert abs(geo.atoms[1].position.y - (-0.388 / 0.529177210)) < 1e-14
    assert abs(geo.atoms[0].position.x - (0.751 / 0.529177210)) < 1e-14
