"""Resolution of the identity in the Coulomb metric on the B200 backend (mirror of ResolutionIdentityModule.jl).

computeMatrixResolutionOfTheIdentityCoulombMetric(basis, basis_aux)   -- :41-50   B[mu,nu,P] = (mu nu|P)(P|Q)^-1/2
computeTensorElectronRepulsionIntegralsRICoulomb(basis, basis_aux)    -- :54-58   B B^T as an N^4 array (tests only)
computeMatrixExchangeRIK(basis, basis_aux, density)                   -- :69-78   K without ever forming N^4
computeMatrixCoulombRIJ(basis, basis_aux, density)                    -- new: the matching RI-J

`basis` / `basis_aux` are shell lists (or a B200Shells for `basis`; a GaussianBasis is regrouped into shells); the
auxiliary shells (up to f) go through the same normalisation contract as orbital shells.  B200RI keeps B on the device
(packed over mu >= nu, this rank's slice of the auxiliary index) for repeated builds inside an SCF loop;
`build_jk_occ(Cocc)` is the occupied-orbital form an SCF driver uses (P = Cocc Cocc').
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import libqlb
from .b200shells import B200Shells
from .shell import flattenShells, totalBasisFunctions


def _as_shell_list(basis):
    from .basisfunctions import GaussianBasis, shells_from_functions

    if isinstance(basis, GaussianBasis):
        shells, shell_of, comp_of, scale = shells_from_functions(basis.contractedBFs)
        from .shell import nbf

        if len(basis.contractedBFs) != sum(nbf(s) for s in shells) or not np.allclose(scale, 1.0, rtol=1e-12):
            raise ValueError("the RI path needs a GaussianBasis made of complete shells (convert(GaussianBasis, shells))")
        return shells
    return basis


class B200RI:
    def __init__(self, basis, basis_aux):
        basis, basis_aux = _as_shell_list(basis), _as_shell_list(basis_aux)
        self._own = not isinstance(basis, B200Shells)
        self.orbital = B200Shells(basis) if self._own else basis
        self._arrays = flattenShells(list(basis_aux))
        centers, L, nprim, exps, coefs = self._arrays
        h = C.c_void_p()
        libqlb.check(libqlb.load().qlb_ri_create(self.orbital.handle, len(basis_aux), libqlb.dptr(centers), libqlb.iptr(L),
                                                 libqlb.iptr(nprim), libqlb.dptr(exps), libqlb.dptr(coefs), C.byref(h)))
        self._h = h
        self.nbf = self.orbital.nbf
        self.naux = totalBasisFunctions(list(basis_aux))
        self.last_ms = None

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            libqlb.load().qlb_ri_destroy(self._h)
            self._h = None
            if self._own:
                self.orbital.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def b_tensor(self) -> np.ndarray:
        out = np.zeros((self.nbf, self.nbf, self.naux), dtype=np.float64, order="F")
        libqlb.check(libqlb.load().qlb_ri_b_tensor(self._h, libqlb.dptr(out)))
        return out

    def eri_tensor(self) -> np.ndarray:
        n = self.nbf
        out = np.empty((n, n, n, n), dtype=np.float64, order="F")
        libqlb.check(libqlb.load().qlb_ri_eri_tensor(self._h, libqlb.dptr(out)))
        return out

    def build_jk(self, density, want_j=True, want_k=True):
        P = np.asfortranarray(density, dtype=np.float64)
        if P.shape != (self.nbf, self.nbf):
            raise ValueError(f"density matrix must be {self.nbf}x{self.nbf}")
        J = np.empty_like(P, order="F") if want_j else None
        K = np.empty_like(P, order="F") if want_k else None
        ms = C.c_double()
        libqlb.check(libqlb.load().qlb_ri_build_jk(self._h, libqlb.dptr(P), libqlb.dptr(J) if want_j else None,
                                                   libqlb.dptr(K) if want_k else None, C.byref(ms)))
        self.last_ms = ms.value
        return J, K

    def build_jk_occ(self, Cocc, want_j=True, want_k=True):
        """J and K for P = Cocc Cocc' from the occupied orbital coefficients (N x nocc)."""
        Cm = np.asfortranarray(Cocc, dtype=np.float64)
        if Cm.ndim != 2 or Cm.shape[0] != self.nbf:
            raise ValueError(f"Cocc must be {self.nbf} x nocc")
        J = np.empty((self.nbf, self.nbf), order="F") if want_j else None
        K = np.empty((self.nbf, self.nbf), order="F") if want_k else None
        ms = C.c_double()
        libqlb.check(libqlb.load().qlb_ri_build_jk_occ(self._h, libqlb.dptr(Cm), Cm.shape[1], libqlb.dptr(J) if want_j else None,
                                                       libqlb.dptr(K) if want_k else None, C.byref(ms)))
        self.last_ms = ms.value
        return J, K

    def slice(self):
        q0, nq = C.c_int32(), C.c_int32()
        libqlb.check(libqlb.load().qlb_ri_slice(self._h, C.byref(q0), C.byref(nq)))
        return q0.value, nq.value


def _with_ri(basis, basis_aux, fn):
    ri = B200RI(basis, basis_aux)
    try:
        return fn(ri)
    finally:
        ri.destroy()


def computeMatrixResolutionOfTheIdentityCoulombMetric(basis, basis_aux) -> np.ndarray:
    return _with_ri(basis, basis_aux, lambda ri: ri.b_tensor())


computeMatrixResolutionOfTheIdentity = computeMatrixResolutionOfTheIdentityCoulombMetric


def computeTensorElectronRepulsionIntegralsRICoulomb(basis, basis_aux) -> np.ndarray:
    return _with_ri(basis, basis_aux, lambda ri: ri.eri_tensor())


def computeMatrixExchangeRIK(basis, basis_aux, density) -> np.ndarray:
    return _with_ri(basis, basis_aux, lambda ri: ri.build_jk(density, want_j=False)[1])


def computeMatrixCoulombRIJ(basis, basis_aux, density) -> np.ndarray:
    return _with_ri(basis, basis_aux, lambda ri: ri.build_jk(density, want_k=False)[0])
