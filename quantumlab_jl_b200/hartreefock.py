"""Restricted Hartree-Fock driver (mirror of HartreeFockModule.jl): computeEnergyHartreeFock (:14-71),
evaluateSCFStep (:73-84), evaluateSCF primitive form (:104-140) and convenience form (:158-208).

Iteration semantics are the reference's: plain Roothaan steps, no DIIS/damping, J and K evaluated on the initial
guess before the loop, convergence on |dE| < energyConvergenceCriterion; the returned energy excludes the
interatomic repulsion (:37,40).  The generalised eigenproblem stays on the host (LAPACK), as in the reference.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .b200shells import B200Shells
from .base import Ionization
from .basisset import BasisSet
from .geometry import Geometry, computeEnergyInteratomicRepulsion, computeNumberElectrons
from .initialguess import computeDensityMatrixGuess
from .shell import computeBasisShells
from .specialmatrices import computeMatrixCoulomb, computeMatrixExchange


def computeEnergyHartreeFock(density, kinetic, nuclearAttraction, coulomb, exchange, interatomicRepulsion, info: bool = True) -> float:
    P = density
    eT = 2 * np.trace(kinetic @ P)
    eV = 2 * np.trace(nuclearAttraction @ P)
    eJ = 2 * np.trace(coulomb @ P)
    eK = -np.trace(exchange @ P)
    if info:
        print("=================================")
        print(f"Kinetic Energy: {eT}")
        print(f"Nuclear Attraction Energy: {eV}")
        print(f"Sum 1-Electron Energy: {eT + eV}")
        print("---------------------------------")
        print(f"Coulomb Energy: {eJ}")
        print(f"Exchange Energy: {eK}")
        print(f"Total Electronic Energy: {eT + eV + eJ + eK}")
        print(f"Total Energy: {eT + eV + eJ + eK + interatomicRepulsion}")
        print("=================================")
    return eT + eV + eJ + eK


def evaluateSCFStep(fock, overlap, electronNumber: int):
    eigenvalues, eigenvectors = scipy.linalg.eigh(fock, overlap)
    Cocc = eigenvectors[:, :electronNumber]
    return eigenvalues, Cocc @ Cocc.T


def _evaluateSCF_primitive(P, S, T, V, coulomb, exchange, Vnn, nocc, info=True, detailedinfo=None,
                           energyConvergenceCriterion=1e-8, maxIterations=500):
    detailedinfo = info if detailedinfo is None else detailedinfo
    totalEnergy = computeEnergyHartreeFock(P, T, V, coulomb(P), exchange(P), Vnn, info=False)
    energies = np.zeros(0)
    J, K = coulomb(P), exchange(P)
    i = 0
    while True:
        i += 1
        F = T + V + 2 * J - K
        energies, P = evaluateSCFStep(F, S, nocc)
        J, K = coulomb(P), exchange(P)
        oldEnergy = totalEnergy
        totalEnergy = computeEnergyHartreeFock(P, T, V, J, K, Vnn, info=detailedinfo)
        if info:
            print(f"E[{i}]: {totalEnergy}")
        if abs(totalEnergy - oldEnergy) < energyConvergenceCriterion:
            if info:
                print("Converged!")
            break
        if i >= maxIterations:
            raise RuntimeError("evaluateSCF did not converge")
    evaluateSCF.last_iterations = i
    return energies, totalEnergy, P


_SCF_KWARGS = {"info", "detailedinfo", "energyConvergenceCriterion", "maxIterations", "computeTensorElectronRepulsionIntegralsCoulomb",
               "computeTensorElectronRepulsionIntegralsExchange", "deviceSCF", "incremental", "basisSetName"}


def evaluateSCF(*args, **kw):
    """evaluateSCF(Pinit, S, T, V, coulomb, exchange, Vnn, nocc; ...)   -- HartreeFockModule.jl:104-140
    evaluateSCF(basis, geometry, initialGuess, electronNumber=Ionization(0); ...)  -- :158-208
    `basis` is a B200Shells, a list of Shell, a GaussianBasis or a BasisSet.  Keyword arguments are the reference's
    (info, detailedinfo, energyConvergenceCriterion, computeTensorElectronRepulsionIntegralsCoulomb/Exchange: providers
    of a stored N^4 tensor that replace the direct J / K build, :166-167, :200-204) plus
      deviceSCF (default True)   keep P, J, K, F and the eigenproblem on the device (qlb_scf_*)
      incremental                density-difference Fock builds (8f-3)
      maxIterations, basisSetName (needed by SADGuess).
    J and K always come from ONE fused device pass per iteration."""
    unknown = set(kw) - _SCF_KWARGS
    if unknown:
        raise TypeError(f"evaluateSCF: unknown keyword argument(s) {sorted(unknown)}")
    if len(args) == 8:
        prim = {k: kw[k] for k in ("info", "detailedinfo", "energyConvergenceCriterion", "maxIterations") if k in kw}
        return _evaluateSCF_primitive(*args, **prim)
    if len(args) not in (3, 4):
        raise TypeError("evaluateSCF takes (Pinit,S,T,V,coulomb,exchange,Vnn,nocc) or (basis,geometry,initialGuess[,electronNumber])")
    from .basisfunctions import FunctionMap, GaussianBasis
    from .devicescf import evaluateSCF_device
    from .initialguess import SADGuess
    from .specialmatrices import (computeMatrixKinetic, computeMatrixNuclearAttraction, computeMatrixOverlap)

    basis, geometry, guess = args[:3]
    electronNumber = args[3] if len(args) > 3 else Ionization(0)
    if isinstance(geometry, str):
        geometry = Geometry.from_file(geometry)
    own = False
    basis_name = kw.get("basisSetName")
    if isinstance(basis, BasisSet):
        basis_name = basis_name or getattr(basis, "name", None)
        basis = computeBasisShells(basis, geometry)
    info = kw.get("info", True)
    loop = dict(info=info, detailedinfo=kw.get("detailedinfo", info),
                energyConvergenceCriterion=kw.get("energyConvergenceCriterion", 1e-8), maxIterations=kw.get("maxIterations", 500))
    if isinstance(electronNumber, Ionization):
        electronNumber = computeNumberElectrons(geometry, electronNumber.charge) // 2
    tensorC = kw.get("computeTensorElectronRepulsionIntegralsCoulomb")
    tensorX = kw.get("computeTensorElectronRepulsionIntegralsExchange")

    def initial_density(n):
        if guess is SADGuess:
            if basis_name is None:
                raise ValueError("SADGuess needs the basis-set name: pass a BasisSet or basisSetName=...")
            return computeDensityMatrixGuess(guess, n, basis_name, geometry)
        return computeDensityMatrixGuess(guess, n)

    if isinstance(basis, GaussianBasis) or tensorC is not None or tensorX is not None:
        # function-list basis and / or stored-tensor providers: the closure form of the loop (host eigenproblem),
        # every J / K still evaluated on the device
        S = computeMatrixOverlap(basis)
        T = computeMatrixKinetic(basis)
        V = computeMatrixNuclearAttraction(basis, geometry)
        Vnn = computeEnergyInteratomicRepulsion(geometry)
        carg = tensorC(basis) if tensorC is not None else basis
        xarg = tensorX(basis) if tensorX is not None else basis
        if isinstance(basis, GaussianBasis) and tensorC is None and tensorX is None:
            fmap = FunctionMap(basis.contractedBFs)
            dev = B200Shells(fmap.shells)
            try:
                def jk(P, which, _cache={}):
                    key = P.tobytes()
                    if _cache.get("key") != key:
                        J, K = dev.build_jk(fmap.embed_density(P))
                        _cache.update(key=key, J=fmap.extract_matrix(J), K=fmap.extract_matrix(K))
                    return _cache[which]
                return _evaluateSCF_primitive(initial_density(S.shape[0]), S, T, V, lambda P: jk(P, "J"), lambda P: jk(P, "K"),
                                              Vnn, electronNumber, **loop)
            finally:
                dev.destroy()
        return _evaluateSCF_primitive(initial_density(S.shape[0]), S, T, V, lambda P: computeMatrixCoulomb(carg, P),
                                      lambda P: computeMatrixExchange(xarg, P), Vnn, electronNumber, **loop)
    if not isinstance(basis, B200Shells):
        basis, own = B200Shells(basis), True
    try:
        S = basis.one_electron("overlap")
        T = basis.one_electron("kinetic")
        V = basis.one_electron("nuclear", geometry)
        Pinit = initial_density(S.shape[0])
        Vnn = computeEnergyInteratomicRepulsion(geometry)
        if kw.get("deviceSCF", True):
            out = evaluateSCF_device(basis, Pinit, S, T, V, Vnn, electronNumber, incremental=kw.get("incremental", basis.incremental), **loop)
            evaluateSCF.last_iterations = evaluateSCF_device.last_iterations
            return out
        basis.incremental = bool(kw.get("incremental", basis.incremental))
        return _evaluateSCF_primitive(Pinit, S, T, V, lambda P: computeMatrixCoulomb(basis, P),
                                      lambda P: computeMatrixExchange(basis, P), Vnn, electronNumber, **loop)
    finally:
        if own:
            basis.destroy()


evaluateSCF.last_iterations = 0
