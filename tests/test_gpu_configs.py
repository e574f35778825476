"""GPU parity at the FULL size of every BASELINE.json configuration (tests/golden/make_golden_configs.py fixtures):
J and K against the oracle's golden columns and whole-matrix checksums, bit-exact screening counts (oracle rule on the
device's own Schwarz bounds), energies of the first Roothaan steps, the sharded deal (round-robin and measured-cost
schedule) summed on one device, the device-resident SCF loop and the library's DMMA GEMM."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import atoms_of, make_basis

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAU = 1e-12


def _synthetic_density(shells, seed=7):  # the generator's density (make_golden_configs.synthetic_density)
    cen = np.array([s.center for s in shells for _ in range((s.lqn.exponent + 1) * (s.lqn.exponent + 2) // 2)])
    d = np.linalg.norm(cen[:, None, :] - cen[None, :, :], axis=2)
    rng = np.random.default_rng(seed)
    A = rng.standard_normal(d.shape) * 0.2 * np.exp(-0.35 * d)
    return A + A.T


@pytest.fixture(scope="module")
def qlb():
    from quantumlab_jl_b200 import libqlb

    libqlb.init()
    return libqlb


@pytest.mark.parametrize("name", ["benzene_ccpvdz", "h2o32_631gs", "c40h82_def2svp", "h2o64_ccpvdz"])
def test_full_size_jk_against_golden(orc, qlb, name):
    from quantumlab_jl_b200.b200shells import B200Shells

    path = os.path.join(GOLD, name + "_jk.npz")
    if not os.path.isfile(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    geo, shells, b, nocc = make_basis(orc, name)
    P = _synthetic_density(shells)
    dev = B200Shells(shells)
    J, K = dev.build_jk(P, tau=TAU)
    st = dev.last_stats
    cols = g["cols"]
    assert np.abs(J[:, cols] - g["Jc"]).max() < 1e-10 and np.abs(K[:, cols] - g["Kc"]).max() < 1e-10
    r = np.random.default_rng(11).standard_normal(b.N)
    # every element of J and K enters these: |r| ~ 1, so 1e-10 per element allows ~sqrt(N) 1e-10 here
    assert np.abs(J @ r - g["Jr"]).max() < 1e-9 and np.abs(K @ r - g["Kr"]).max() < 1e-9
    assert abs(J.sum() - float(g["sumJ"])) < 1e-7 and abs(K.sum() - float(g["sumK"])) < 1e-7
    assert abs(np.trace(J @ P) - float(g["trJP"])) < 1e-8 and abs(np.trace(K @ P) - float(g["trKP"])) < 1e-8
    assert np.array_equal(J, J.T) and np.array_equal(K, K.T)
    # screening count: bit-exact against the oracle's integer rule fed the DEVICE's Schwarz bounds
    kept, unique = orc.screen_count(b, dev.schwarz(), P, tau=TAU)
    assert st["quartets_unique"] == unique == int(g["counts"][1])
    assert st["quartets_kept"] == kept
    # the oracle's own bounds agree with the device's to ~1e-13: the two counts can only differ at bin edges
    assert abs(kept - int(g["counts"][0])) <= 1e-5 * kept
    dev.destroy()


@pytest.mark.parametrize("name", ["benzene_ccpvdz", "h2o32_631gs"])
def test_energy_after_roothaan_steps(orc, qlb, name):
    """E (electronic + Vnn) after k = 1..3 Roothaan steps from the core-Hamiltonian guess == oracle-driven SCF, 1e-8."""
    from quantumlab_jl_b200.b200shells import B200Shells
    from quantumlab_jl_b200.geometry import computeEnergyInteratomicRepulsion
    from quantumlab_jl_b200.hartreefock import computeEnergyHartreeFock, evaluateSCFStep

    path = os.path.join(GOLD, name + "_scf.npz")
    if not os.path.isfile(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    geo, shells, b, nocc = make_basis(orc, name)
    dev = B200Shells(shells)
    S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
    Vnn = computeEnergyInteratomicRepulsion(geo)
    assert abs(Vnn - float(g["Vnn"])) < 1e-9
    eps, P = evaluateSCFStep(T + V, S, nocc)
    for k, Eref in enumerate(g["E"]):
        J, K = dev.build_jk(P, tau=TAU)
        E = computeEnergyHartreeFock(P, T, V, J, K, Vnn, info=False) + Vnn
        # relative 1e-11 of |E| ~ 1e3 Eh is the FP64 floor of the traces; the bar is 1e-8 Eh
        assert abs(E - Eref) < 1e-8, (k, E, Eref)
        eps, P = evaluateSCFStep(T + V + 2 * J - K, S, nocc)
    dev.destroy()


def _device_build(qlb, dev, P, rank, nranks, stats=False):
    import torch

    lib = qlb.load()
    n = dev.nbf
    dP = torch.tensor(np.asfortranarray(P).T.copy(), dtype=torch.float64, device="cuda")  # column-major == transposed row-major
    dJK = torch.zeros(2 * n * n, dtype=torch.float64, device="cuda")
    st = qlb.Stats()
    qlb.check(lib.qlb_build_jk_device(dev.handle, C.c_void_p(dP.data_ptr()), TAU, C.c_void_p(dJK.data_ptr()),
                                      C.c_void_p(dJK.data_ptr() + 8 * n * n), rank, nranks, None, C.byref(st)))
    torch.cuda.synchronize()
    return dJK.cpu().numpy(), st.as_dict()


@pytest.mark.parametrize("name,nranks", [("h2o4_631gs", 2), ("h2o4_631gs", 3), ("benzene_ccpvdz", 8), ("c2h6_def2svp", 3)])
def test_sharded_pieces_sum_to_full_build(orc, qlb, name, nranks):
    """The library's own multi-rank deal, evaluated piece by piece on ONE device: the partial [J;K] of ranks 0..n-1 sum
    to the single-rank build, and the kept-quartet counts add up exactly -- for the round-robin deal of the first
    (calibration) build and for the measured-cost schedule of every later build."""
    from quantumlab_jl_b200.b200shells import B200Shells

    geo, shells, b, nocc = make_basis(orc, name)
    P = _synthetic_density(shells, seed=5)
    dev = B200Shells(shells)
    lib = qlb.load()
    lib.qlb_set_profiling(1)
    full, st_full = _device_build(qlb, dev, P, 0, 1, stats=True)
    lib.qlb_set_profiling(0)
    # (1) round-robin deal: a fresh handle, so that the first sharded build is the calibration build
    rr = B200Shells(shells)
    acc = np.zeros_like(full)
    kept = 0
    for r in range(nranks):
        part, st = _device_build(qlb, rr, P, r, nranks)
        acc += part
        kept += st["quartets_kept"]
    assert np.abs(acc - full).max() < 1e-12 and kept == st_full["quartets_kept"]
    rr.destroy()
    # (2) measured-cost schedule: feed the feedback a real multi-rank run would have all-reduced
    counts = np.zeros(2 * qlb.NQCLASS)
    ms = np.zeros(qlb.NQCLASS)
    for i, cname in enumerate(qlb.QCLASS_NAMES):
        counts[2 * i] = st_full["quartets_kept_class"].get(cname, 0)
        counts[2 * i + 1] = st_full["prim_quartets_class"].get(cname, 0)
        ms[i] = st_full["ms_class"].get(cname, 0.0)
    lib.qlb_debug_set_schedule_feedback.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    qlb.check(lib.qlb_debug_set_schedule_feedback(dev.handle, nranks, qlb.dptr(counts), qlb.dptr(ms)))
    acc = np.zeros_like(full)
    kept = 0
    for r in range(nranks):
        qlb.check(lib.qlb_debug_set_schedule_feedback(dev.handle, nranks, qlb.dptr(counts), qlb.dptr(ms)))
        part, st = _device_build(qlb, dev, P, r, nranks)
        acc += part
        kept += st["quartets_kept"]
    assert np.abs(acc - full).max() < 1e-12 and kept == st_full["quartets_kept"]
    # the schedule the library used is the one qlb_schedule_pieces reports: every piece has exactly one owner
    cost = ms.copy()
    npieces = np.zeros(qlb.NQCLASS, dtype=np.int32)
    owner = np.zeros(qlb.NQCLASS * nranks, dtype=np.int32)
    qlb.check(lib.qlb_schedule_pieces(qlb.dptr(cost), nranks, qlb.iptr(npieces), qlb.iptr(owner)))
    for qc in range(qlb.NQCLASS):
        own = owner[qc * nranks:(qc + 1) * nranks]
        assert np.all(own[:npieces[qc]] >= 0) and np.all(own[npieces[qc]:] == -1)
    dev.destroy()


def test_device_scf_h2o_sto3g_reference_kat(orc, qlb):
    """runtests.jl:84 with the whole loop on the device: -74.96178985 Eh, 11 Roothaan iterations from ZeroGuess."""
    from quantumlab_jl_b200.b200shells import B200Shells
    from quantumlab_jl_b200.devicescf import evaluateSCF_device
    from quantumlab_jl_b200.geometry import computeEnergyInteratomicRepulsion

    geo, shells, b, nocc = make_basis(orc, "h2o_sto3g")
    dev = B200Shells(shells)
    S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
    Vnn = computeEnergyInteratomicRepulsion(geo)
    eps, E, P = evaluateSCF_device(dev, np.zeros_like(S), S, T, V, Vnn, nocc, info=False)
    assert abs(E + Vnn - (-74.96178985)) < 1e-7
    assert evaluateSCF_device.last_iterations == 11
    assert abs(np.trace(P @ S) - nocc) < 1e-10 and np.array_equal(P, P.T)
    dev.destroy()


@pytest.mark.parametrize("name", ["h2o_631gs", "benzene_ccpvdz"])
def test_device_scf_matches_host_loop(orc, qlb, name):
    """Same iteration count and energy (1e-8) as the host loop (scipy eigh), d shells included; also incremental builds."""
    from quantumlab_jl_b200.b200shells import B200Shells
    from quantumlab_jl_b200.devicescf import evaluateSCF_device
    from quantumlab_jl_b200.geometry import computeEnergyInteratomicRepulsion
    from quantumlab_jl_b200.hartreefock import evaluateSCF, evaluateSCFStep
    from quantumlab_jl_b200.specialmatrices import computeMatrixCoulomb, computeMatrixExchange

    geo, shells, b, nocc = make_basis(orc, name)
    dev = B200Shells(shells)
    S = dev.one_electron("overlap"); T = dev.one_electron("kinetic"); V = dev.one_electron("nuclear", geo)
    Vnn = computeEnergyInteratomicRepulsion(geo)
    _, P0 = evaluateSCFStep(T + V, S, nocc)  # core-Hamiltonian start: plain Roothaan from zero oscillates on benzene
    maxit = 60
    try:
        eps_h, E_h, P_h = evaluateSCF(P0, S, T, V, lambda P: computeMatrixCoulomb(dev, P), lambda P: computeMatrixExchange(dev, P),
                                      Vnn, nocc, info=False, maxIterations=maxit)
        it_h = evaluateSCF.last_iterations
    except RuntimeError:
        pytest.skip("plain Roothaan iterations do not converge for this system (the reference has no damping either)")
    eps_d, E_d, P_d = evaluateSCF_device(dev, P0, S, T, V, Vnn, nocc, info=False, maxIterations=maxit)
    assert evaluateSCF_device.last_iterations == it_h
    assert abs(E_d - E_h) < 1e-8
    assert np.abs(P_d - P_h).max() < 1e-6 and np.abs(eps_d - eps_h).max() < 1e-7
    eps_i, E_i, P_i = evaluateSCF_device(dev, P0, S, T, V, Vnn, nocc, info=False, maxIterations=maxit, incremental=True)
    assert abs(E_i - E_h) < 1e-8
    dev.destroy()


@pytest.mark.parametrize("ta,tb,M,N,K", [(0, 0, 64, 64, 16), (0, 1, 200, 200, 37), (1, 0, 130, 67, 301), (1, 1, 33, 129, 5),
                                         (0, 1, 608, 608, 160), (0, 0, 1, 1, 1),
                                         # the 128 x 128 / 128 x 64 tile kernels, every layout pair, ragged edges
                                         (0, 0, 256, 192, 100), (0, 1, 300, 130, 77), (1, 0, 130, 260, 301), (1, 1, 257, 129, 50),
                                         (0, 0, 400, 160, 33), (1, 1, 128, 64, 16),
                                         # split-K (few tiles, long contraction index)
                                         (0, 1, 130, 130, 4000), (1, 0, 64, 60, 3000), (0, 0, 300, 200, 1111)])
def test_dmma_gemm_matches_numpy(qlb, ta, tb, M, N, K):
    import torch

    lib = qlb.load()
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    ref = 0.75 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * C0
    dA = torch.tensor(A.T.copy(), dtype=torch.float64, device="cuda")  # column-major storage
    dB = torch.tensor(B.T.copy(), dtype=torch.float64, device="cuda")
    dC = torch.tensor(C0.T.copy(), dtype=torch.float64, device="cuda")
    qlb.check(lib.qlb_dgemm_device(ta, tb, M, N, K, 0.75, C.c_void_p(dA.data_ptr()), A.shape[0], C.c_void_p(dB.data_ptr()), B.shape[0],
                                   -0.5, C.c_void_p(dC.data_ptr()), M, None))
    torch.cuda.synchronize()
    got = dC.cpu().numpy().T
    assert np.abs(got - ref).max() < 1e-12 * max(1, K)


@pytest.mark.parametrize("N,ncol,nslab", [(20, 7, 3), (200, 130, 3), (150, 64, 2), (131, 131, 1)])
def test_dmma_packed_symmetric_products(qlb, N, ncol, nslab):
    """LAY_SYM operand of the DMMA kernels (small and large tiles): X^q = B^q R and K = sum_q X^q B^q against numpy."""
    import torch

    lib = qlb.load()
    rng = np.random.default_rng(N + ncol)
    Bs = rng.standard_normal((nslab, N, N))
    Bs = Bs + Bs.transpose(0, 2, 1)
    packed = np.concatenate([np.concatenate([Bs[q][b:, b] for b in range(N)]) for q in range(nslab)])
    R = rng.standard_normal((N, ncol))
    dB = torch.tensor(packed, dtype=torch.float64, device="cuda")
    dR = torch.tensor(R.T.copy(), dtype=torch.float64, device="cuda")
    dX = torch.zeros(nslab * ncol * N, dtype=torch.float64, device="cuda")
    qlb.check(lib.qlb_dsymm_packed_device(0, N, ncol, nslab, C.c_void_p(dB.data_ptr()), C.c_void_p(dR.data_ptr()), C.c_void_p(dX.data_ptr()), None))
    torch.cuda.synchronize()
    X = dX.cpu().numpy().reshape(nslab, ncol, N).transpose(0, 2, 1)
    assert np.abs(X - Bs @ R).max() < 1e-12 * N
    Y = rng.standard_normal((nslab, N, N))
    dY = torch.tensor(Y.transpose(0, 2, 1).copy(), dtype=torch.float64, device="cuda")
    dK = torch.zeros(N * N, dtype=torch.float64, device="cuda")
    qlb.check(lib.qlb_dsymm_packed_device(1, N, N, nslab, C.c_void_p(dB.data_ptr()), C.c_void_p(dY.data_ptr()), C.c_void_p(dK.data_ptr()), None))
    torch.cuda.synchronize()
    K = dK.cpu().numpy().reshape(N, N).T
    assert np.abs(K - sum(Y[q] @ Bs[q] for q in range(nslab))).max() < 1e-12 * N * nslab


def test_single_process_multi_gpu_matches_single_gpu(orc, qlb):
    """qlb_multi_*: one process drives every visible GPU; the sharded + all-reduced J, K equal the one-GPU build."""
    import torch

    from quantumlab_jl_b200.b200shells import B200MultiShells, B200Shells

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    geo, shells, b, nocc = make_basis(orc, "benzene_ccpvdz")
    P = _synthetic_density(shells, seed=3)
    one = B200Shells(shells)
    J1, K1 = one.build_jk(P, tau=TAU)
    multi = B200MultiShells(shells, ngpu)
    for it in range(4):  # calibration build, first scheduled build, steady state
        J, K = multi.build_jk(P, tau=TAU)
        assert np.abs(J - J1).max() < 1e-12 and np.abs(K - K1).max() < 1e-12, f"build {it}"
    H = one.one_electron("kinetic") + one.one_electron("nuclear", geo)
    assert np.abs(multi.fock(H, P, tau=TAU) - (H + 2 * J1 - K1)).max() < 1e-11
    multi.destroy()
    one.destroy()


def test_function_list_entry_points(orc, qlb):
    """a3 / a6 / a12 / a16: the basis-function flavours of the reference API on the device path -- the scalar
    computeElectronRepulsionIntegral over four primitive / four contracted functions (IntegralsModule.jl:412-462), the
    CGBF-vector block form (:464-470), GaussianBasis tensor, J/K and one-electron matrices, and the stored-tensor flavour of
    J/K (SpecialMatricesModule.jl:227-229, 264-266) contracted on the device."""
    from quantumlab_jl_b200.basisfunctions import GaussianBasis, PrimitiveGaussianBasisFunction, computeBasis
    from quantumlab_jl_b200.basisset import loadBasisSet
    from quantumlab_jl_b200.base import MQuantumNumber
    from quantumlab_jl_b200.integrals import (computeElectronRepulsionIntegral, computeTensorBlockElectronRepulsionIntegrals,
                                              computeTensorElectronRepulsionIntegrals)
    from quantumlab_jl_b200.specialmatrices import (computeMatrixCoulomb, computeMatrixExchange, computeMatrixFock,
                                                    computeMatrixOverlap)

    geo, shells, b, nocc = make_basis(orc, "h2o_631gs")
    ref = orc.eri_tensor(b, literal=False)  # Cook closed form, exact Boys, functions = expandShell of the shells
    bas = GaussianBasis.from_shells(shells)  # convert(GaussianBasis, shells)
    T = computeTensorElectronRepulsionIntegrals(bas)
    assert np.abs(T - ref).max() < 1e-10
    f = bas.contractedBFs
    for (i, j, k, l) in [(0, 1, 2, 3), (13, 7, 5, 18), (16, 16, 14, 2), (4, 9, 9, 4)]:  # s, p and d functions (d_xy included)
        assert abs(computeElectronRepulsionIntegral(f[i], f[j], f[k], f[l]) - ref[i, j, k, l]) < 1e-10
    blk = computeTensorBlockElectronRepulsionIntegrals(f[13:19], f[0:1], f[2:5], f[13:19])  # CGBF-vector form
    assert np.abs(blk - ref[13:19, 0:1, 2:5, 13:19]).max() < 1e-10
    # four primitives: (mu nu|la si) of un-normalised primitives == oracle's primitive Cook formula
    pr = [PrimitiveGaussianBasisFunction(shells[1].center, 1.3, MQuantumNumber(1, 0, 0)),
          PrimitiveGaussianBasisFunction(shells[0].center, 0.7, MQuantumNumber(0, 0, 0)),
          PrimitiveGaussianBasisFunction(shells[1].center, 0.9, MQuantumNumber(0, 1, 1)),
          PrimitiveGaussianBasisFunction(shells[-1].center, 2.1, MQuantumNumber(1, 0, 0))]
    want = orc.eri_prim([p.center for p in pr], [p.exponent for p in pr], [tuple(p.mqn) for p in pr], mode=orc.BOYS_EXACT, literal=False)
    assert abs(computeElectronRepulsionIntegral(*pr) - want) < 1e-10 * max(1.0, abs(want))
    # GaussianBasis flavour of J / K / S; computeBasis normalises every function by its own self-overlap
    P = _synthetic_density(shells, seed=9)
    J, K = orc.coulomb_from_tensor(ref, P), orc.exchange_from_tensor(ref, P)
    assert np.abs(computeMatrixCoulomb(bas, P) - J).max() < 1e-10 and np.abs(computeMatrixExchange(bas, P) - K).max() < 1e-10
    bas2 = computeBasis(loadBasisSet("6-31g*"), geo)
    S2 = computeMatrixOverlap(bas2)
    assert np.abs(np.diag(S2) - 1.0).max() < 1e-12
    assert np.abs(computeMatrixCoulomb(bas2, P) - J).max() < 1e-10  # the two normalisation routes agree
    # stored-tensor flavour: contracted on the device, not with numpy
    assert np.abs(computeMatrixCoulomb(T, P) - J).max() < 1e-11 and np.abs(computeMatrixExchange(T, P) - K).max() < 1e-11
    F3 = computeMatrixFock(shells, geo, P)
    F5 = computeMatrixFock(shells, geo, P, T, T)
    assert np.abs(F3 - F5).max() < 1e-9
