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


def evaluateSCF(*args, **kw):
    """evaluateSCF(Pinit, S, T, V, coulomb, exchange, Vnn, nocc; ...)   -- HartreeFockModule.jl:104-140
    evaluateSCF(basis, geometry, initialGuess, electronNumber=Ionization(0); ...)  -- :158-208
    `basis` is a B200Shells, a list of Shell or a BasisSet; J and K come from ONE fused device pass per iteration."""
    if len(args) == 8:
        return _evaluateSCF_primitive(*args, **kw)
    basis, geometry, guess = args[:3]
    electronNumber = args[3] if len(args) > 3 else Ionization(0)
    if isinstance(geometry, str):
        geometry = Geometry.from_file(geometry)
    own = False
    if isinstance(basis, BasisSet):
        basis = computeBasisShells(basis, geometry)
    if not isinstance(basis, B200Shells):
        basis, own = B200Shells(basis), True
    try:
        if isinstance(electronNumber, Ionization):
            electronNumber = computeNumberElectrons(geometry, electronNumber.charge) // 2
        S = basis.one_electron("overlap")
        T = basis.one_electron("kinetic")
        V = basis.one_electron("nuclear", geometry)
        Pinit = computeDensityMatrixGuess(guess, S.shape[0])
        Vnn = computeEnergyInteratomicRepulsion(geometry)
        return _evaluateSCF_primitive(Pinit, S, T, V, lambda P: computeMatrixCoulomb(basis, P),
                                      lambda P: computeMatrixExchange(basis, P), Vnn, electronNumber,
                                      info=kw.get("info", True), detailedinfo=kw.get("detailedinfo", kw.get("info", True)),
                                      energyConvergenceCriterion=kw.get("energyConvergenceCriterion", 1e-8))
    finally:
        if own:
            basis.destroy()


evaluateSCF.last_iterations = 0
