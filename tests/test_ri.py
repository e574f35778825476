"""Resolution of the identity (ResolutionIdentityModule.jl): oracle pinned against the reference's known answers
(test/runtests.jl:138-141), CUDA path against the oracle."""
import numpy as np
import pytest

from conftest import make_basis


def _sad_alpha(geo):
    from quantumlab_jl_b200.initialguess import SAD_STO3G_ALPHA, sad_density

    return sad_density(geo, SAD_STO3G_ALPHA)


def test_oracle_ri_reference_kats(orc):
    geo, shells, b, _ = make_basis(orc, "h2o_sto3g")
    B = orc.ri_b_tensor(b, b)  # the reference's tests use the orbital basis as auxiliary basis
    T = np.einsum("mnq,lsq->mnls", B, B)
    assert abs(T.mean() - 0.0429373705056905) < 1e-8         # runtests.jl:140
    assert abs(T[0, 2, 3, 4] - 0.0012436634924295) < 1e-9    # runtests.jl:141  (1-based [1,3,4,5])
    K = orc.ri_exchange(B, _sad_alpha(geo))
    assert abs(K.mean() - 0.3950752513027109) < 1e-8         # runtests.jl:138


@pytest.mark.gpu
def test_cuda_ri_reference_kats(orc):
    from quantumlab_jl_b200.resolutionidentity import (computeMatrixExchangeRIK,
                                                       computeTensorElectronRepulsionIntegralsRICoulomb)

    geo, shells, b, _ = make_basis(orc, "h2o_sto3g")
    T = computeTensorElectronRepulsionIntegralsRICoulomb(shells, shells)
    assert abs(T.mean() - 0.0429373705056905) < 1e-8
    assert abs(T[0, 2, 3, 4] - 0.0012436634924295) < 1e-9
    K = computeMatrixExchangeRIK(shells, shells, _sad_alpha(geo))
    assert abs(K.mean() - 0.3950752513027109) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h2o_sto3g", "h2o_631gs", "c2h6_def2svp"])
def test_cuda_ri_against_oracle(orc, name):
    from quantumlab_jl_b200.resolutionidentity import B200RI

    geo, shells, b, _ = make_basis(orc, name)
    ri = B200RI(shells, shells)  # aux = orbital shells: s, p (and d) auxiliary functions, every (ab|P) class
    B = ri.b_tensor()
    Br = orc.ri_b_tensor(b, b)
    # B is defined up to the conditioning of (P|Q)^-1/2: compare the fitted integrals and the contractions
    T, Tr = np.einsum("mnq,lsq->mnls", B, B), np.einsum("mnq,lsq->mnls", Br, Br)
    assert np.abs(T - Tr).max() < 1e-8
    rng = np.random.default_rng(3)
    A = rng.standard_normal((b.N, b.N)) * 0.3
    P = A + A.T
    J, K = ri.build_jk(P)
    assert np.abs(J - orc.ri_coulomb(Br, P)).max() < 1e-8
    assert np.abs(K - orc.ri_exchange(Br, P)).max() < 1e-8
    # RI with aux = orbital product space is not exact, but the diagonal (mu nu|mu nu) fit is variational: below exact
    ri.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h2o_631gs", "h2o2_631gs"])
def test_cuda_ri_jkfit_aux_with_f_shells(orc, name):
    """A JK-fitting auxiliary basis with f shells (data/basis/jkfit-et-dz.tx93): B against the oracle's numpy RI, the
    general-P and the occupied-orbital contractions against each other and the oracle, the fit against the exact J/K."""
    from quantumlab_jl_b200 import workloads
    from quantumlab_jl_b200.b200shells import B200Shells
    from quantumlab_jl_b200.hartreefock import evaluateSCFStep
    from quantumlab_jl_b200.resolutionidentity import B200RI
    from quantumlab_jl_b200.shell import flattenShells

    geo, shells, b, nocc = make_basis(orc, name)
    aux = workloads.jkfit_aux(geo)
    assert max(s.lqn.exponent for s in aux) == 3
    baux = orc.Basis(*flattenShells(aux))
    dev = B200Shells(shells)
    ri = B200RI(dev, aux)
    assert ri.slice() == (0, baux.N)
    Br = orc.ri_b_tensor(b, baux)
    B = ri.b_tensor()
    T, Tr = np.einsum("mnq,lsq->mnls", B, B), np.einsum("mnq,lsq->mnls", Br, Br)
    assert np.abs(T - Tr).max() < 1e-8
    S = dev.one_electron("overlap"); H = dev.one_electron("kinetic") + dev.one_electron("nuclear", geo)
    import scipy.linalg

    w, v = scipy.linalg.eigh(H, S)
    Cocc = np.asfortranarray(v[:, :nocc])
    P = Cocc @ Cocc.T
    J, K = ri.build_jk(P)
    Jo, Ko = ri.build_jk_occ(Cocc)
    assert np.abs(J - orc.ri_coulomb(Br, P)).max() < 1e-8 and np.abs(K - orc.ri_exchange(Br, P)).max() < 1e-8
    assert np.abs(Jo - J).max() < 1e-10 and np.abs(Ko - K).max() < 1e-10
    assert np.array_equal(Ko, Ko.T)
    # quality of the stand-in auxiliary basis: RI-J/K energies within 1e-3 relative of the exact ones
    Jx, Kx = dev.build_jk(P, tau=0.0)
    assert abs(np.trace(J @ P) / np.trace(Jx @ P) - 1) < 1e-3 and abs(np.trace(K @ P) / np.trace(Kx @ P) - 1) < 1e-3
    ri.destroy()
    dev.destroy()
