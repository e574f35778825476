"""B200Shells -- the shell-vector type that selects the B200 backend.

The reference dispatches `computeMatrixCoulomb/Exchange/Fock`, `computeTensor...` and `evaluateSCF` on the type of
their basis argument: Vector{Shell} (native Julia), Vector{LibInt2Shell} (libint2 engine; LibInt2Module.jl:62-68).
`B200Shells` is the third member of that family: a list of `Shell`s plus the opaque device handle (`qlb_basis`)
that owns the device-resident shells, primitive-pair records, Schwarz bounds and class bins.  Like LibInt2Shell it
is destroyed explicitly (`destroy()`; LibInt2Module.jl:108-110) or by the garbage collector.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import libqlb
from .shell import Shell, flattenShells, totalBasisFunctions

DEFAULT_TAU = 1e-12  # Schwarz x density threshold (protects the 1e-10 per-element parity)
MAX_L = 2            # QLB_MAX_L of include/qlb200.h: orbital shells up to d (the RI auxiliary index goes to f)


def _check_orbital_L(shells):
    """The class kernels exist for s, p and d orbital shells; name the offender instead of a bare C error code."""
    for i, s in enumerate(shells):
        if s.lqn.exponent > MAX_L:
            raise ValueError(f"shell {i} (L = {s.lqn.exponent}, center {tuple(s.center)}) exceeds the orbital angular momentum "
                             f"this library is built for (QLB_MAX_L = {MAX_L}: s, p, d); f and higher shells are only supported "
                             f"on the RI auxiliary index")


class B200Shells(list):
    def __init__(self, shells, tau: float = DEFAULT_TAU, device: int = -1):
        super().__init__(shells)
        if not all(isinstance(s, Shell) for s in self):
            raise TypeError("B200Shells takes a sequence of Shell")
        _check_orbital_L(self)
        lib = libqlb.init(device)
        self._arrays = flattenShells(list(self))  # keep alive
        centers, L, nprim, exps, coefs = self._arrays
        h = C.c_void_p()
        libqlb.check(lib.qlb_basis_create(len(self), libqlb.dptr(centers), libqlb.iptr(L), libqlb.iptr(nprim),
                                          libqlb.dptr(exps), libqlb.dptr(coefs), C.byref(h)))
        self._h = h
        self.tau = float(tau)
        self.nbf = totalBasisFunctions(list(self))
        self.last_stats: dict | None = None
        self._cache_P = None
        self._cache_JK = None
        self.incremental = False      # density-difference builds inside an SCF loop
        self.rebuild_every = 8
        self._since_full = 0
        self.build_history: list = []  # quartets kept by each build issued through jk_cached

    # ---- handle management
    @property
    def handle(self):
        if self._h is None:
            raise libqlb.QlbError("B200Shells handle was destroyed")
        return self._h

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            libqlb.load().qlb_basis_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- hot path
    def _check_P(self, P) -> np.ndarray:
        P = np.asfortranarray(P, dtype=np.float64)
        if P.shape != (self.nbf, self.nbf):
            raise ValueError(f"density matrix must be {self.nbf}x{self.nbf}, got {P.shape}")
        # the kernels read one triangle of P and symmetrise J and K afterwards (every SCF density is symmetric); the
        # reference's nsh^4 loop also accepts a non-symmetric matrix -- refuse it instead of answering for its symmetric part
        scale = float(np.abs(P).max()) if P.size else 0.0
        if scale > 0.0 and float(np.abs(P - P.T).max()) > 1e-10 * scale:
            raise ValueError("density matrix must be symmetric (the device path digests one triangle of P)")
        return P

    def build_jk(self, P, tau: float | None = None, want_stats: bool = True):
        """One fused, screened pass -> (J, K) exactly as the reference defines them."""
        P = self._check_P(P)
        n = self.nbf
        J = np.empty((n, n), dtype=np.float64, order="F")
        K = np.empty((n, n), dtype=np.float64, order="F")
        st = libqlb.Stats()
        libqlb.check(libqlb.load().qlb_build_jk(self.handle, libqlb.dptr(P), self.tau if tau is None else tau,
                                                libqlb.dptr(J), libqlb.dptr(K), C.byref(st) if want_stats else None))
        if want_stats:
            self.last_stats = st.as_dict()
        return J, K

    def jk_cached(self, P):
        """J and K for P, computed together once: `coulomb(P)` followed by `exchange(P)` costs one pass.

        With `incremental=True` a new density is handled as J(P) = J(P_prev) + J(P - P_prev): J and K
        are linear in P and the density DIFFERENCE shrinks as the SCF converges, so the Schwarz x density rule keeps
        fewer and fewer quartets.  Every `rebuild_every`-th build is a full one, which bounds the accumulated
        screening error (each build drops contributions below tau)."""
        P = self._check_P(P)
        if self._cache_P is not None and self._cache_P.shape == P.shape and np.array_equal(self._cache_P, P):
            return self._cache_JK
        if self.incremental and self._cache_P is not None and self._since_full < self.rebuild_every and np.any(self._cache_P):
            dJ, dK = self.build_jk(P - self._cache_P)
            self._cache_JK = (self._cache_JK[0] + dJ, self._cache_JK[1] + dK)
            self._since_full += 1
        else:
            self._cache_JK = self.build_jk(P)
            self._since_full = 0
        self._cache_P = P.copy()
        self.build_history.append(self.last_stats["quartets_kept"] if self.last_stats else None)
        return self._cache_JK

    def fock(self, H, P, tau: float | None = None):
        P = self._check_P(P)
        H = np.asfortranarray(H, dtype=np.float64)
        F = np.empty_like(P, order="F")
        st = libqlb.Stats()
        libqlb.check(libqlb.load().qlb_fock(self.handle, libqlb.dptr(H), libqlb.dptr(P), self.tau if tau is None else tau,
                                            libqlb.dptr(F), C.byref(st)))
        self.last_stats = st.as_dict()
        return F

    def schwarz(self) -> np.ndarray:
        Q = np.zeros((len(self), len(self)), dtype=np.float64, order="F")
        libqlb.check(libqlb.load().qlb_schwarz(self.handle, libqlb.dptr(Q)))
        return Q

    def eri_tensor(self) -> np.ndarray:
        n = self.nbf
        out = np.empty((n, n, n, n), dtype=np.float64, order="F")
        libqlb.check(libqlb.load().qlb_eri_tensor(self.handle, libqlb.dptr(out)))
        return out

    def eri_block(self, i: int, j: int, k: int, l: int) -> np.ndarray:
        from .shell import nbf

        if not all(0 <= s < len(self) for s in (i, j, k, l)):
            raise libqlb.QlbError(f"shell index out of range 0..{len(self) - 1}: {(i, j, k, l)}")
        shp = tuple(nbf(self[s]) for s in (i, j, k, l))
        out = np.empty(shp, dtype=np.float64, order="F")
        libqlb.check(libqlb.load().qlb_eri_block(self.handle, i, j, k, l, libqlb.dptr(out)))
        return out

    def one_electron(self, which: str, geometry=None) -> np.ndarray:
        n = self.nbf
        out = np.empty((n, n), dtype=np.float64, order="F")
        lib = libqlb.load()
        if which == "overlap":
            libqlb.check(lib.qlb_overlap(self.handle, libqlb.dptr(out)))
        elif which == "kinetic":
            libqlb.check(lib.qlb_kinetic(self.handle, libqlb.dptr(out)))
        elif which == "nuclear":
            xyz = np.ascontiguousarray([a.position for a in geometry.atoms], dtype=np.float64).reshape(-1)
            Z = np.ascontiguousarray([a.element.atomicNumber for a in geometry.atoms], dtype=np.float64)
            libqlb.check(lib.qlb_nuclear_attraction(self.handle, Z.size, libqlb.dptr(xyz), libqlb.dptr(Z), libqlb.dptr(out)))
        else:
            raise ValueError(which)
        return out


class B200MultiShells(list):
    """One process, several GPUs (include/qlb200.h: qlb_multi_*): the shell-vector type a single evaluateSCF caller uses
    to put the whole box behind `computeMatrixCoulomb/Exchange/Fock`.  Every device holds the basis; a build is sharded
    over the devices and combined by one grouped all-reduce; device 0 answers."""

    def __init__(self, shells, ngpu: int, tau: float = DEFAULT_TAU):
        super().__init__(shells)
        _check_orbital_L(self)
        lib = libqlb.load()
        self._arrays = flattenShells(list(self))
        centers, L, nprim, exps, coefs = self._arrays
        h = C.c_void_p()
        libqlb.check(lib.qlb_multi_create(int(ngpu), len(self), libqlb.dptr(centers), libqlb.iptr(L), libqlb.iptr(nprim),
                                          libqlb.dptr(exps), libqlb.dptr(coefs), C.byref(h)))
        libqlb._initialised = True
        self._h = h
        self.ngpu = int(ngpu)
        self.tau = float(tau)
        self.nbf = totalBasisFunctions(list(self))
        self.last_stats = None

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            libqlb.load().qlb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def build_jk(self, P, tau: float | None = None):
        P = np.asfortranarray(P, dtype=np.float64)
        if P.shape != (self.nbf, self.nbf):
            raise ValueError(f"density matrix must be {self.nbf}x{self.nbf}, got {P.shape}")
        J = np.empty_like(P, order="F")
        K = np.empty_like(P, order="F")
        st = libqlb.Stats()
        libqlb.check(libqlb.load().qlb_multi_build_jk(self._h, libqlb.dptr(P), self.tau if tau is None else tau, libqlb.dptr(J),
                                                      libqlb.dptr(K), C.byref(st)))
        self.last_stats = st.as_dict()  # device 0's share
        return J, K

    def fock(self, H, P, tau: float | None = None):
        P = np.asfortranarray(P, dtype=np.float64)
        H = np.asfortranarray(H, dtype=np.float64)
        F = np.empty_like(P, order="F")
        libqlb.check(libqlb.load().qlb_multi_fock(self._h, libqlb.dptr(H), libqlb.dptr(P), self.tau if tau is None else tau,
                                                  libqlb.dptr(F), None))
        return F
