"""GPU parity tests: libqlb200 (through its C ABI, via the host-side mirror of the reference API) against the CPU
oracle on the same inputs.  Tolerances are the north star's: 1e-10 Eh per integral and per Fock element, 1e-8 Eh on
the RHF energy; screening counts bit-exact."""
import numpy as np
import pytest

from conftest import atoms_of, make_basis

pytestmark = pytest.mark.gpu

TOL_INT = 1e-10


@pytest.fixture(scope="module")
def qlb():
    from quantumlab_jl_b200 import libqlb

    libqlb.init()
    return libqlb


def _b200(shells, **kw):
    from quantumlab_jl_b200.b200shells import B200Shells

    return B200Shells(shells, **kw)


def _rand_sym(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) * scale
    return A + A.T


def test_device_is_blackwell(qlb):
    info = qlb.device_info()
    assert info["cc"][0] >= 10 and info["sm_count"] > 0


def test_eri_tensor_h2o_sto3g(orc, qlb):
    geo, shells, b, _ = make_basis(orc, "h2o_sto3g")
    ref = orc.eri_tensor(b, literal=True)
    dev = _b200(shells)
    got = dev.eri_tensor()
    assert np.abs(got - ref).max() < 1e-12
    # reference KAT runtests.jl:120 and the BASELINE.md golden
    blk = dev.eri_block(0, 4, 3, 3)
    assert abs(blk[0, 0, 1, 1] - got[0, 6, 4, 4]) < 1e-14
    assert abs(blk[0, 0, 1, 1] - 0.1492228667247951) < 1e-12
    # the literal reference (Gamma-difference Boys, clamp) agrees for s/p shells to its own accuracy
    quirk = orc.eri_tensor(b, mode=orc.BOYS_REFQUIRK, literal=True)
    assert np.abs(got - quirk).max() < 1e-8  # the Gamma-difference Boys routine is only ~1e-9 accurate here


@pytest.mark.parametrize("name", ["h2o_631gs", "h2o_ccpvdz", "c2h6_def2svp"])
def test_eri_blocks_all_classes(orc, qlb, name):
    geo, shells, b, _ = make_basis(orc, name)
    dev = _b200(shells)
    rng = np.random.default_rng(3)
    byL = {}
    for i, s in enumerate(shells):
        byL.setdefault(s.lqn.exponent, []).append(i)
    Ls = sorted(byL)
    quartets = []
    for la in Ls:
        for lb in Ls:
            for lc in Ls:
                for ld in Ls:
                    quartets.append(tuple(int(rng.choice(byL[l])) for l in (la, lb, lc, ld)))
    worst = 0.0
    for (i, j, k, l) in quartets:
        ref = orc.eri_block(b, i, j, k, l, literal=False)  # Cook closed form, exact Boys: the oracle of record
        got = dev.eri_block(i, j, k, l)
        assert got.shape == ref.shape
        worst = max(worst, np.abs(got - ref).max())
    assert worst < TOL_INT, worst


def test_schwarz_bounds(orc, qlb):
    geo, shells, b, _ = make_basis(orc, "h2o2_631gs")
    dev = _b200(shells)
    Q = dev.schwarz()
    Qr = orc.schwarz(b)
    assert np.abs(Q - Qr).max() < 1e-12 * max(1.0, Qr.max())
    assert np.allclose(Q, Q.T)


@pytest.mark.parametrize("name,tau", [("h2o_sto3g", 0.0), ("h2o_631gs", 0.0), ("h2o2_631gs", 1e-12), ("h2o4_631gs", 1e-10),
                                      ("c2h6_def2svp", 1e-12), ("h2o_ccpvdz", 1e-12)])
def test_jk_against_oracle(orc, qlb, name, tau):
    geo, shells, b, _ = make_basis(orc, name)
    dev = _b200(shells)
    P = _rand_sym(b.N, 7, 0.3)
    J, K = dev.build_jk(P, tau=tau)
    Q = dev.schwarz()
    Jr, Kr, counts = orc.md_jk(b, P, tau=tau if tau > 0 else 1e-300, Q=Q)
    assert np.abs(J - Jr).max() < TOL_INT and np.abs(K - Kr).max() < TOL_INT
    assert np.array_equal(J, J.T) and np.array_equal(K, K.T)
    st = dev.last_stats
    assert st["quartets_unique"] == counts[1]
    assert st["quartets_kept"] == counts[0]  # bit-exact screening count (integer rule, device bounds)
    F = dev.fock(np.zeros_like(P), P, tau=tau)
    assert np.abs(F - (2 * Jr - Kr)).max() < 3 * TOL_INT


def test_jk_equals_reference_full_loop(orc, qlb):
    # the reference's own contraction: all nsh^4 ordered quartets, separate J and K passes
    geo, shells, b, _ = make_basis(orc, "h2o_sto3g")
    dev = _b200(shells)
    P = _rand_sym(b.N, 11)
    Jr, _ = orc.ref_contract(b, P, "J", literal=True)
    Kr, _ = orc.ref_contract(b, P, "K", literal=True)
    J, K = dev.build_jk(P, tau=0.0)
    assert np.abs(J - Jr).max() < TOL_INT and np.abs(K - Kr).max() < TOL_INT


def test_screening_counts_density_dependent(orc, qlb):
    geo, shells, b, _ = make_basis(orc, "h2o4_631gs")
    dev = _b200(shells)
    Q = dev.schwarz()
    rng = np.random.default_rng(5)
    for scale, tau in ((1.0, 1e-8), (1e-3, 1e-9), (1e-6, 1e-10)):
        A = rng.standard_normal((b.N, b.N)) * scale * np.exp(-3 * rng.random((b.N, b.N)) * 8)
        P = A + A.T
        dev.build_jk(P, tau=tau)
        kept, vis = orc.screen_count(b, Q, P, tau)
        assert dev.last_stats["quartets_kept"] == kept and dev.last_stats["quartets_unique"] == vis
        assert 0 < kept < vis


def test_jk_linearity_and_zero(orc, qlb):
    geo, shells, b, _ = make_basis(orc, "h2o2_631gs")
    dev = _b200(shells)
    P1, P2 = _rand_sym(b.N, 1), _rand_sym(b.N, 2)
    J1, K1 = dev.build_jk(P1, tau=0.0)
    J2, K2 = dev.build_jk(P2, tau=0.0)
    J3, K3 = dev.build_jk(P1 + 2 * P2, tau=0.0)
    assert np.abs(J3 - (J1 + 2 * J2)).max() < 1e-9 and np.abs(K3 - (K1 + 2 * K2)).max() < 1e-9
    J0, K0 = dev.build_jk(np.zeros((b.N, b.N)))
    assert not J0.any() and not K0.any()
    assert dev.last_stats["quartets_kept"] == 0


def test_one_electron_matrices(orc, qlb):
    for name in ("h2o_sto3g", "h2o_631gs", "c2h6_def2svp"):
        geo, shells, b, _ = make_basis(orc, name)
        xyz, Z = atoms_of(geo)
        dev = _b200(shells)
        assert np.abs(dev.one_electron("overlap") - orc.one_electron(b, "overlap")).max() < 1e-12
        assert np.abs(dev.one_electron("kinetic") - orc.one_electron(b, "kinetic")).max() < 1e-11
        assert np.abs(dev.one_electron("nuclear", geo) - orc.one_electron(b, "nuclear", xyz, Z)).max() < 1e-10


def test_scf_h2o_sto3g_reference_kat(orc, qlb):
    # runtests.jl:84: total RHF energy -74.96178985 (1e-7); BASELINE.md: 11 Roothaan iterations from ZeroGuess
    from quantumlab_jl_b200.geometry import computeEnergyInteratomicRepulsion
    from quantumlab_jl_b200.hartreefock import evaluateSCF
    from quantumlab_jl_b200.initialguess import ZeroGuess

    geo, shells, b, nocc = make_basis(orc, "h2o_sto3g")
    eps, E, P = evaluateSCF(_b200(shells), geo, ZeroGuess, nocc, info=False)
    assert evaluateSCF.last_iterations == 11
    assert abs(E + computeEnergyInteratomicRepulsion(geo) - (-74.96178985)) < 1e-7
    assert abs(E - (-84.2254524222)) < 1e-8
    # same through the Ionization default and a plain Shell list
    eps2, E2, P2 = evaluateSCF(shells, geo, ZeroGuess, info=False)
    assert abs(E2 - E) < 1e-12


def test_fock_sad_kat(orc, qlb):
    # runtests.jl:148 mean(computeMatrixFock(bas,h2o,SAD alpha)) = -0.785008186026 (1e-7)
    from quantumlab_jl_b200.initialguess import SAD_STO3G_ALPHA, sad_density
    from quantumlab_jl_b200.specialmatrices import computeMatrixFock

    geo, shells, b, _ = make_basis(orc, "h2o_sto3g")
    F = computeMatrixFock(shells, geo, sad_density(geo, SAD_STO3G_ALPHA))
    assert abs(F.mean() - (-0.785008186026)) < 1e-7


def test_scf_energy_dshell_vs_oracle(orc, qlb):
    # d shells: oracle of record = closed form with exact Boys; SCF energy to 1e-8 Eh
    from quantumlab_jl_b200.hartreefock import evaluateSCF
    from quantumlab_jl_b200.initialguess import ZeroGuess

    geo, shells, b, nocc = make_basis(orc, "h2o_631gs")
    xyz, Z = atoms_of(geo)
    S, T = orc.one_electron(b, "overlap"), orc.one_electron(b, "kinetic")
    V = orc.one_electron(b, "nuclear", xyz, Z)
    Q = orc.schwarz(b)
    jk = {}

    def oj(P):
        key = P.tobytes()
        if key not in jk:
            jk.clear()
            jk[key] = orc.md_jk(b, P, tau=1e-14, Q=Q)
        return jk[key]

    _, Er, Pr, itr = orc.evaluate_scf(np.zeros((b.N, b.N)), S, T, V, lambda P: oj(P)[0], lambda P: oj(P)[1], nocc)
    eps, E, P = evaluateSCF(_b200(shells), geo, ZeroGuess, nocc, info=False)
    assert abs(E - Er) < 1e-8
    assert evaluateSCF.last_iterations == itr


def test_incremental_fock_builds(orc, qlb):
    # density-difference builds reach the same SCF energy with far fewer quartets in late iterations
    from quantumlab_jl_b200.hartreefock import evaluateSCF
    from quantumlab_jl_b200.initialguess import ZeroGuess

    geo, shells, b, nocc = make_basis(orc, "h2o2_631gs")
    full = _b200(shells)
    _, E_full, _ = evaluateSCF(full, geo, ZeroGuess, nocc, info=False)
    it_full = evaluateSCF.last_iterations
    inc = _b200(shells)
    inc.incremental = True
    _, E_inc, _ = evaluateSCF(inc, geo, ZeroGuess, nocc, info=False)
    assert evaluateSCF.last_iterations == it_full
    assert abs(E_inc - E_full) < 1e-8
    kept_full = [k for k in full.build_history if k]
    kept_inc = [k for k in inc.build_history if k]
    assert min(kept_inc) < 0.95 * min(kept_full)  # the converging density difference is screened harder (small, dense system)


def test_benzene_ccpvdz_full_build(orc, qlb):
    geo, shells, b, nocc = make_basis(orc, "benzene_ccpvdz")
    assert b.N == 120 and b.nsh == 54
    dev = _b200(shells)
    P = _rand_sym(b.N, 13, 0.2)
    J, K = dev.build_jk(P, tau=1e-12)
    Jr, Kr, counts = orc.md_jk(b, P, tau=1e-12, Q=dev.schwarz())
    assert dev.last_stats["quartets_kept"] == counts[0]
    assert np.abs(J - Jr).max() < TOL_INT and np.abs(K - Kr).max() < TOL_INT


def test_bad_arguments_raise(qlb):
    from quantumlab_jl_b200.base import LQuantumNumber, origin
    from quantumlab_jl_b200.shell import Shell

    with pytest.raises(ValueError, match="QLB_MAX_L"):  # the host mirror names the offending shell ...
        _b200([Shell(origin, LQuantumNumber("F"), [1.0], [1.0])])  # L > QLB_MAX_L, cf. LibInt2Module.jl:79-83
    import ctypes as C

    lib = qlb.init(-1)
    c, L_, n, e, k = (np.zeros(3), np.array([3], dtype=np.int32), np.array([1], dtype=np.int32), np.ones(1), np.ones(1))
    h = C.c_void_p()
    with pytest.raises(qlb.QlbError):                   # ... and the C ABI rejects it on its own
        qlb.check(lib.qlb_basis_create(1, qlb.dptr(c), qlb.iptr(L_), qlb.iptr(n), qlb.dptr(e), qlb.dptr(k), C.byref(h)))
    dev = _b200([Shell(origin, LQuantumNumber("S"), [1.0], [1.0])])
    with pytest.raises(ValueError):
        dev.build_jk(np.zeros((3, 3)))
    two = _b200([Shell(origin, LQuantumNumber("S"), [1.0], [1.0]), Shell(origin, LQuantumNumber("S"), [0.3], [1.0])])
    with pytest.raises(ValueError, match="symmetric"):
        two.build_jk(np.array([[1.0, 0.2], [0.1, 1.0]]))
    two.destroy()
    with pytest.raises(qlb.QlbError):
        dev.eri_block(0, 0, 0, 5)
    dev.destroy()
    with pytest.raises(qlb.QlbError):
        dev.build_jk(np.zeros((1, 1)))


def _basis_from(orc, shells):
    from quantumlab_jl_b200.shell import flattenShells

    return orc.Basis(*flattenShells(shells))


def test_edge_cases_tiny_ragged_and_distant(orc, qlb):
    from quantumlab_jl_b200.base import LQuantumNumber, Position
    from quantumlab_jl_b200.shell import Shell

    S, P_, D = LQuantumNumber("S"), LQuantumNumber("P"), LQuantumNumber("D")
    cases = {
        "one s shell": [Shell(Position(0, 0, 0), S, [0.8], [1.0])],
        "one d shell": [Shell(Position(0.1, -0.2, 0.3), D, [1.1, 0.3], [0.6, 0.5])],
        "ragged contractions": [Shell(Position(0, 0, 0), S, [50.0, 8.0, 2.0, 0.6, 0.2], [0.05, 0.2, 0.4, 0.4, 0.2]),
                                Shell(Position(0, 0, 1.4), P_, [0.9], [1.0]),
                                Shell(Position(1.0, 0.5, 0), D, [0.7], [1.0]),
                                Shell(Position(0, 0, 1.4), S, [0.3, 0.1], [0.7, 0.4])],
        "two distant atoms": [Shell(Position(0, 0, 0), S, [1.2, 0.3], [0.5, 0.6]), Shell(Position(0, 0, 0), P_, [0.5], [1.0]),
                              Shell(Position(0, 0, 60.0), S, [1.2, 0.3], [0.5, 0.6]), Shell(Position(0, 0, 60.0), P_, [0.5], [1.0])],
    }
    for name, shells in cases.items():
        b = _basis_from(orc, shells)
        dev = _b200(shells)
        P = _rand_sym(b.N, 21, 0.5)
        for tau in (0.0, 1e-12):
            J, K = dev.build_jk(P, tau=tau)
            Jr, Kr, counts = orc.md_jk(b, P, tau=tau if tau > 0 else 1e-300, Q=dev.schwarz())
            assert np.abs(J - Jr).max() < TOL_INT and np.abs(K - Kr).max() < TOL_INT, name
            if tau > 0:  # tau = 0 switches screening off on the device; the oracle still drops exact-zero bounds
                assert dev.last_stats["quartets_kept"] == counts[0], name
        if name == "two distant atoms":
            assert dev.last_stats["quartets_kept"] < dev.last_stats["quartets_unique"]  # cross-centre pairs are screened out
        T = dev.eri_tensor()
        assert np.abs(T - orc.eri_tensor(b, literal=False)).max() < TOL_INT, name
        # everything screened: nothing launched, exact zeros
        J0, K0 = dev.build_jk(P, tau=1e30)
        assert not J0.any() and not K0.any() and dev.last_stats["quartets_kept"] == 0
        dev.destroy()


def test_build_after_zero_density_is_complete(orc, qlb):
    """The grids of a build are sized from the PREVIOUS build's density bound: a build on the zero density (ZeroGuess) or on a
    tiny density must not make the next build lose quartets; and a 1000x larger density after a small one neither."""
    geo, shells, b, _ = make_basis(orc, "h2o4_631gs")
    dev = _b200(shells)
    P = _rand_sym(b.N, 11, 0.3)
    J0, K0 = dev.build_jk(P, tau=1e-12)
    kept0 = dev.last_stats["quartets_kept"]
    for prev in (np.zeros_like(P), 1e-9 * P, 1e3 * P):
        dev.build_jk(prev, tau=1e-12)
        J, K = dev.build_jk(P, tau=1e-12)
        assert dev.last_stats["quartets_kept"] == kept0
        assert np.abs(J - J0).max() < 1e-12 and np.abs(K - K0).max() < 1e-12
    dev.destroy()
