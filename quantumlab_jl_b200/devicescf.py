"""Device-resident SCF state: the host mirror of `qlb_scf` (include/qlb200.h).

`evaluateSCF` (HartreeFockModule.jl:104-140) per iteration calls evaluateSCFStep (:73-84: generalised eigenproblem,
P = Cocc*Cocc'), the J/K closures and computeEnergyHartreeFock (:14-41).  With a DeviceSCF all of that happens in HBM
(cuSOLVER Dsygvd, the library's DMMA GEMM for the density, one fused reduction for the four traces) and only the four
energy components cross to the host per iteration.  The iteration / convergence rule stays the reference's.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import libqlb
from .b200shells import B200Shells

_ITEMS = {"P": 0, "F": 1, "J": 2, "K": 3, "eps": 4, "C": 5}


class DeviceSCF:
    def __init__(self, basis: B200Shells, S, T, V, nocc: int):
        self.basis = basis
        self.n = basis.nbf
        self.nocc = int(nocc)
        lib = libqlb.load()
        S, T, V = (np.asfortranarray(M, dtype=np.float64) for M in (S, T, V))
        for M in (S, T, V):
            if M.shape != (self.n, self.n):
                raise ValueError(f"S, T, V must be {self.n}x{self.n}")
        h = C.c_void_p()
        libqlb.check(lib.qlb_scf_create(basis.handle, libqlb.dptr(S), libqlb.dptr(T), libqlb.dptr(V), self.nocc, C.byref(h)))
        self._h = h
        self.last_stats = None

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            libqlb.load().qlb_scf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def set_density(self, P):
        P = np.asfortranarray(P, dtype=np.float64)
        if P.shape != (self.n, self.n):
            raise ValueError(f"density matrix must be {self.n}x{self.n}")
        libqlb.check(libqlb.load().qlb_scf_set_density(self._h, libqlb.dptr(P)))

    def build(self, tau: float | None = None, incremental: bool = False, want_stats: bool = False):
        """J, K <- P on the device; -> (eT, eV, eJ, eK) = (2 tr TP, 2 tr VP, 2 tr JP, -tr KP)."""
        e = np.zeros(4)
        st = libqlb.Stats()
        libqlb.check(libqlb.load().qlb_scf_build(self._h, self.basis.tau if tau is None else tau, int(incremental), libqlb.dptr(e),
                                                 C.byref(st) if want_stats else None))
        if want_stats:
            self.last_stats = st.as_dict()
        return tuple(e)

    def step(self):
        """F = T + V + 2J - K; eigen(F, S); P = Cocc Cocc' (evaluateSCFStep)."""
        libqlb.check(libqlb.load().qlb_scf_step(self._h))

    def get(self, what: str) -> np.ndarray:
        n = self.n
        out = np.empty(n if what == "eps" else (n, n), dtype=np.float64, order="F")
        libqlb.check(libqlb.load().qlb_scf_get(self._h, _ITEMS[what], libqlb.dptr(out)))
        return out


def evaluateSCF_device(basis: B200Shells, Pinit, S, T, V, Vnn: float, nocc: int, info: bool = True, detailedinfo=None,
                       energyConvergenceCriterion: float = 1e-8, maxIterations: int = 500, incremental: bool = False):
    """HartreeFockModule.jl:104-140 with every matrix resident on the device.  Returns (orbital energies, electronic
    energy, P) like the reference; `evaluateSCF_device.last_iterations` holds the iteration count."""
    detailedinfo = info if detailedinfo is None else detailedinfo
    scf = DeviceSCF(basis, S, T, V, nocc)
    try:
        scf.set_density(Pinit)
        # the reference evaluates J and K on the initial guess (twice, :118 and :123-124); one build gives the same numbers
        total = sum(scf.build(incremental=False, want_stats=True))
        basis.build_history.append(scf.last_stats["quartets_kept"])
        i = 0
        while True:
            i += 1
            scf.step()
            eT, eV, eJ, eK = scf.build(incremental=incremental, want_stats=True)
            basis.build_history.append(scf.last_stats["quartets_kept"])
            old, total = total, eT + eV + eJ + eK
            if detailedinfo:
                print("=================================")
                print(f"Kinetic Energy: {eT}")
                print(f"Nuclear Attraction Energy: {eV}")
                print(f"Sum 1-Electron Energy: {eT + eV}")
                print("---------------------------------")
                print(f"Coulomb Energy: {eJ}")
                print(f"Exchange Energy: {eK}")
                print(f"Total Electronic Energy: {total}")
                print(f"Total Energy: {total + Vnn}")
                print("=================================")
            if info:
                st = scf.last_stats or {}
                ms = st.get("ms_total", 0.0)
                kept = st.get("quartets_kept", 0)
                print(f"E[{i}]: {total}   (Fock build {ms:.2f} ms, {kept} quartets kept, "
                      f"{kept / ms * 1e-6 if ms else 0.0:.1f} M quartets/s)")
            if abs(total - old) < energyConvergenceCriterion:
                if info:
                    print("Converged!")
                break
            if i >= maxIterations:
                raise RuntimeError("evaluateSCF did not converge")
        evaluateSCF_device.last_iterations = i
        return scf.get("eps"), total, scf.get("P")
    finally:
        scf.destroy()


evaluateSCF_device.last_iterations = 0
