"""ctypes wrapper + numpy SCF driver around the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  It restates, on top of ql_oracle.c, the reference's matrix-level algorithm:
SpecialMatricesModule.jl:209-290 (J, K, Fock), HartreeFockModule.jl:14-41 (energy), :73-84 (SCF step)
and :104-140 (the SCF loop, including its iteration / convergence rule).

Parity status: pinned by the reference's known-answer tests for H2O/STO-3G (tests/test_oracle_kat.py);
unpinned by the reference for d shells (it has no such test) -- see the header of ql_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

BOYS_EXACT, BOYS_REFQUIRK = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ql_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_boys_exact.restype = C.c_double
        _lib.orc_boys_exact.argtypes = [C.c_int, C.c_double]
        _lib.orc_boys_refquirk.restype = C.c_double
        _lib.orc_boys_refquirk.argtypes = [C.c_int, C.c_double]
        _lib.orc_cook_eri_prim.restype = C.c_double
        _lib.orc_qlog.restype = C.c_int
        _lib.orc_qlog.argtypes = [C.c_double]
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class Basis:
    """Flat SoA description of a shell list: the same five arrays the C ABI takes."""

    def __init__(self, centers, L, nprim, exps, coefs):
        self.centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1)
        self.L = np.ascontiguousarray(L, dtype=np.int32)
        self.nprim = np.ascontiguousarray(nprim, dtype=np.int32)
        self.exps = np.ascontiguousarray(exps, dtype=np.float64)
        self.coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        self.nsh = int(self.L.size)
        self.nbf_sh = (self.L + 1) * (self.L + 2) // 2
        self.boff = np.concatenate([[0], np.cumsum(self.nbf_sh)]).astype(np.int64)
        self.N = int(self.boff[-1])

    def args(self):
        return (C.c_int(self.nsh), _p(self.centers), _p(self.L, C.c_int), _p(self.nprim, C.c_int), _p(self.exps), _p(self.coefs))


def boys(m: int, x: float, mode: int = BOYS_EXACT) -> float:
    return lib().orc_boys_refquirk(m, x) if mode == BOYS_REFQUIRK else lib().orc_boys_exact(m, x)


def shell_renorm(L: int, exps, coefs):
    e = np.ascontiguousarray(exps, dtype=np.float64)
    c = np.array(coefs, dtype=np.float64)
    lib().orc_shell_renorm(C.c_int(L), C.c_int(e.size), _p(e), _p(c))
    return c


def eri_prim(centers, alphas, lmn, mode=BOYS_EXACT, literal=True) -> float:
    c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1)
    a = np.ascontiguousarray(alphas, dtype=np.float64)
    l = np.ascontiguousarray(lmn, dtype=np.int32).reshape(-1)
    return lib().orc_cook_eri_prim(_p(c), _p(a), _p(l, C.c_int), C.c_int(mode), C.c_int(int(literal)))


def eri_block(b: Basis, i, j, k, l, mode=BOYS_EXACT, literal=False, md=False):
    shp = tuple(int(b.nbf_sh[s]) for s in (i, j, k, l))
    out = np.zeros(int(np.prod(shp)), dtype=np.float64)
    if md:
        lib().orc_md_eri_block(*b.args(), C.c_int(i), C.c_int(j), C.c_int(k), C.c_int(l), _p(out))
    else:
        lib().orc_cook_eri_block(*b.args(), C.c_int(i), C.c_int(j), C.c_int(k), C.c_int(l), C.c_int(mode), C.c_int(int(literal)), _p(out))
    return out.reshape(shp, order="F")


def eri_tensor(b: Basis, mode=BOYS_EXACT, literal=False):
    out = np.zeros(b.N**4, dtype=np.float64)
    lib().orc_cook_eri_tensor(*b.args(), C.c_int(mode), C.c_int(int(literal)), _p(out))
    return out.reshape((b.N,) * 4, order="F")


def one_electron(b: Basis, which: str, atoms_xyz=None, atoms_Z=None, mode=BOYS_EXACT):
    w = {"overlap": 0, "kinetic": 1, "nuclear": 2}[which]
    xyz = np.ascontiguousarray(atoms_xyz if atoms_xyz is not None else np.zeros(3), dtype=np.float64).reshape(-1)
    Z = np.ascontiguousarray(atoms_Z if atoms_Z is not None else np.zeros(1), dtype=np.float64)
    nat = 0 if atoms_Z is None else Z.size
    out = np.zeros(b.N * b.N, dtype=np.float64)
    lib().orc_one_electron(*b.args(), C.c_int(w), C.c_int(nat), _p(xyz), _p(Z), C.c_int(mode), _p(out))
    return out.reshape((b.N, b.N), order="F")


def ref_contract(b: Basis, P, which: str, mode=BOYS_EXACT, literal=True, stride=1):
    """The reference's direct J or K: all nsh^4 ordered quartets (SpecialMatricesModule.jl:33-100)."""
    Pf = np.asfortranarray(P, dtype=np.float64)
    out = np.zeros(b.N * b.N, dtype=np.float64)
    nq = C.c_long(0)
    lib().orc_ref_contract(*b.args(), _p(Pf.reshape(-1, order="F")), C.c_int(0 if which == "J" else 1), C.c_int(mode),
                           C.c_int(int(literal)), C.c_long(stride), _p(out), C.byref(nq))
    return out.reshape((b.N, b.N), order="F"), nq.value


def ref_contract_sampled(b: Basis, P, which: str, stride: int, offset: int = 0, nthreads: int = 1, mode=BOYS_REFQUIRK, literal=True):
    """Uniform sample (every stride-th quartet) of the reference's nsh^4 direct contraction; -> (partial matrix, nquartets)."""
    Pf = np.asfortranarray(P, dtype=np.float64)
    out = np.zeros(b.N * b.N, dtype=np.float64)
    f = lib().orc_ref_contract_sampled
    f.restype = C.c_long
    n = f(*b.args(), _p(Pf.reshape(-1, order="F")), C.c_int(0 if which == "J" else 1), C.c_int(mode), C.c_int(int(literal)),
          C.c_long(stride), C.c_long(offset), C.c_int(nthreads), _p(out))
    return out.reshape((b.N, b.N), order="F"), int(n)


def schwarz(b: Basis):
    Q = np.zeros(b.nsh * b.nsh, dtype=np.float64)
    lib().orc_md_schwarz(*b.args(), _p(Q))
    return Q.reshape((b.nsh, b.nsh), order="F")


def md_jk(b: Basis, P, tau=1e-12, Q=None, rank=0, nranks=1):
    Pf = np.asfortranarray(P, dtype=np.float64).reshape(-1, order="F")
    J = np.zeros(b.N * b.N); K = np.zeros(b.N * b.N)
    counts = (C.c_long * 2)()
    qp = _p(np.asfortranarray(Q, dtype=np.float64).reshape(-1, order="F")) if Q is not None else None
    lib().orc_md_jk(*b.args(), _p(Pf), C.c_double(tau), qp, C.c_int(rank), C.c_int(nranks), _p(J), _p(K), counts)
    return J.reshape((b.N, b.N), order="F"), K.reshape((b.N, b.N), order="F"), (counts[0], counts[1])


def screen_count(b: Basis, Q, P, tau=1e-12):
    counts = (C.c_long * 2)()
    Qf = np.asfortranarray(Q, dtype=np.float64).reshape(-1, order="F")
    Pf = np.asfortranarray(P, dtype=np.float64).reshape(-1, order="F")
    lib().orc_screen_count(C.c_int(b.nsh), _p(b.L, C.c_int), _p(Qf), _p(Pf), C.c_double(tau), counts)
    return counts[0], counts[1]


def ri_b_tensor(b_orb: Basis, b_aux: Basis):
    """ResolutionIdentityModule.jl:41-50: B[mu,nu,P] = (mu nu|P 1)(P 1|Q 1)^-1/2 with the unit s function
    (exponent 0, coefficient 1, at the origin) standing in for identityCGF (BasisFunctionsModule.jl:50-51)."""
    centers = np.concatenate([b_orb.centers, b_aux.centers, np.zeros(3)])
    L = np.concatenate([b_orb.L, b_aux.L, [0]]).astype(np.int32)
    nprim = np.concatenate([b_orb.nprim, b_aux.nprim, [1]]).astype(np.int32)
    exps = np.concatenate([b_orb.exps, b_aux.exps, [0.0]])
    coefs = np.concatenate([b_orb.coefs, b_aux.coefs, [1.0]])
    comb = Basis(centers, L, nprim, exps, coefs)
    N, Nx = b_orb.N, b_aux.N
    unit = comb.nsh - 1
    B3 = np.zeros((N, N, Nx))
    for i in range(b_orb.nsh):
        for j in range(b_orb.nsh):
            for k in range(b_aux.nsh):
                blk = eri_block(comb, i, j, b_orb.nsh + k, unit, literal=False)
                B3[b_orb.boff[i]:b_orb.boff[i + 1], b_orb.boff[j]:b_orb.boff[j + 1], b_aux.boff[k]:b_aux.boff[k + 1]] = blk[:, :, :, 0]
    M = np.zeros((Nx, Nx))
    for k in range(b_aux.nsh):
        for l in range(b_aux.nsh):
            blk = eri_block(comb, b_orb.nsh + k, unit, b_orb.nsh + l, unit, literal=False)
            M[b_aux.boff[k]:b_aux.boff[k + 1], b_aux.boff[l]:b_aux.boff[l + 1]] = blk[:, 0, :, 0]
    w, v = np.linalg.eigh(0.5 * (M + M.T))
    Mih = (v * w ** -0.5) @ v.T  # sqrt(inv(C))
    return (B3.reshape(N * N, Nx, order="F") @ Mih).reshape((N, N, Nx), order="F")


def ri_exchange(B, P):  # ResolutionIdentityModule.jl:69-78
    return np.einsum("mnq,ns,lsq->ml", B, P, B, optimize=True)


def ri_coulomb(B, P):
    return np.einsum("mnq,q->mn", B, np.einsum("lsq,ls->q", B, P))


def qlog(x: float) -> int:
    return lib().orc_qlog(x)


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


def num_threads() -> int:
    return lib().orc_num_threads()


# ----------------------------------------------------------------- matrix level (numpy)
def coulomb_from_tensor(eri, P):  # SpecialMatricesModule.jl:227-229
    return np.einsum("mnls,ls->mn", eri, P)


def exchange_from_tensor(eri, P):  # SpecialMatricesModule.jl:264-266
    return np.einsum("mlns,ls->mn", eri, P)


def fock(T, V, J, K):  # SpecialMatricesModule.jl:284-290
    return T + V + 2 * J - K


def energy_hf(P, T, V, J, K):  # HartreeFockModule.jl:14-41 (returned value excludes Vnn)
    return 2 * np.trace(T @ P) + 2 * np.trace(V @ P) + 2 * np.trace(J @ P) - np.trace(K @ P)


def scf_step(F, S, nocc):  # HartreeFockModule.jl:73-84
    import scipy.linalg

    w, v = scipy.linalg.eigh(F, S)
    Cocc = v[:, :nocc]
    return w, Cocc @ Cocc.T


def evaluate_scf(Pinit, S, T, V, coulomb, exchange, nocc, conv=1e-8, maxiter=200):
    """HartreeFockModule.jl:104-140 restated, including the duplicated pre-loop builds."""
    P = Pinit
    total = energy_hf(P, T, V, coulomb(P), exchange(P))
    J, K = coulomb(P), exchange(P)
    it = 0
    eps = None
    while True:
        it += 1
        F = T + V + 2 * J - K
        eps, P = scf_step(F, S, nocc)
        J, K = coulomb(P), exchange(P)
        old, total = total, energy_hf(P, T, V, J, K)
        if abs(total - old) < conv or it >= maxiter:
            break
    return eps, total, P, it
