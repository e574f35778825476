// eri_coop.cuh -- the cooperative flavour of the ERI + J/K class kernels, for the high angular momentum classes
// (total L >= 5) whose Hermite Coulomb integrals R_tuv (56..165 doubles per primitive quartet) do not fit the register
// file next to the accumulators.
//
// Several threads per shell quartet: a CTA is ONE bra shell pair x 32 ket shell pairs (the lanes) x NW = ncart(LC)/CPT
// warps; warp w owns the ket-c Cartesian components w, w+NW, .. (CPT of them), the launch (template parameter DSEL) owns
// ket-d component DSEL, so a thread accumulates CPT*NA*NB integrals (a b | c d_DSEL) of its lane's quartet in registers.
// Per primitive quartet
//   1. every warp evaluates the Boys function and the one-index recursion R^n_{00v} for its lane (redundant, cheap),
//      then builds the (u,v) columns of R_tuv dealt to it entirely in registers and parks them in SHARED memory
//      (layout [t u v][lane]: conflict-free);
//   2. after one barrier each thread contracts R with its own ket Hermite coefficients (static indices: one shared-
//      memory load per R element, "scatter form") into G_tuv, then with the bra coefficients E^ab, which are CTA-uniform
//      and sit in shared memory (broadcast loads), into the accumulators.
// The bra-side Hermite coefficients are built once per CTA (one thread per bra primitive and axis).
// Digestion: the ket-indexed updates go out as FP64 reductions (RED) as in the register flavour; the J_ab contributions
// of the CTA's threads are first summed over the warps (the ket-c components) through shared memory, then over the 32
// lanes with one shuffle tree per element -- per CTA, not per thread (ncu / SASS: a per-thread lane reduction of NA*NB
// values cost more instructions than the integrals of a K = 1 quartet).
// The same kernel serves the J/K build (prm.tensor == nullptr) and the N^4 tensor scatter of the parity tests.
//
// Replaces, like the other flavours, computeElectronRepulsionIntegral (IntegralsModule.jl:412-462) per quartet and
// contractTensorBlocksWithMatrix_* (SpecialMatricesModule.jl:33-100).
#pragma once
#include <type_traits>
#include <utility>

#include "eri_kernels.cuh"

namespace qlb {
namespace cop {

constexpr int COOP_MAXBP = 4;  // bra primitive pairs whose E^ab coefficients are resident in shared memory at a time

__host__ __device__ constexpr int cart_x(int L, int idx) {
  int i = 0;
  for (int x = L; x >= 0; --x)
    for (int y = L - x; y >= 0; --y, ++i)
      if (i == idx) return x;
  return 0;
}
__host__ __device__ constexpr int cart_y(int L, int idx) {
  int i = 0;
  for (int x = L; x >= 0; --x)
    for (int y = L - x; y >= 0; --y, ++i)
      if (i == idx) return y;
  return 0;
}

// (t,u,v) of position idx in the tetrahedral layout hidx(L, ...)
__host__ __device__ constexpr int herm_t(int L, int idx) {
  for (int t = 0; t <= L; ++t)
    for (int u = 0; u <= L - t; ++u)
      for (int v = 0; v <= L - t - u; ++v)
        if (hidx(L, t, u, v) == idx) return t;
  return 0;
}
__host__ __device__ constexpr int herm_u(int L, int idx) {
  for (int t = 0; t <= L; ++t)
    for (int u = 0; u <= L - t; ++u)
      for (int v = 0; v <= L - t - u; ++v)
        if (hidx(L, t, u, v) == idx) return u;
  return 0;
}
__host__ __device__ constexpr int herm_v(int L, int idx) {
  for (int t = 0; t <= L; ++t)
    for (int u = 0; u <= L - t; ++u)
      for (int v = 0; v <= L - t - u; ++v)
        if (hidx(L, t, u, v) == idx) return v;
  return 0;
}

// compile-time loop: f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N-1>)
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>()), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
  static_for_impl(f, std::make_integer_sequence<int, N>());
}

template <int LA, int LB, int LC, int LD, int CPT = 1>
struct CoopShape {
  static_assert(ncart(LC) % CPT == 0, "CPT must divide the number of ket-c components");
  static constexpr int NW = ncart(LC) / CPT;
  static constexpr int L = LA + LB + LC + LD, LAB = LA + LB, LCD = LC + LD;
  static constexpr int NA = ncart(LA), NB = ncart(LB);
  static constexpr int NE1 = (LA + 1) * (LB + 1) * (LAB + 1);  // E^ab entries per axis
  static constexpr int R_DOUBLES = nherm(L) * 32;
  static constexpr int E_DOUBLES = COOP_MAXBP * 3 * NE1;
  static constexpr int B_DOUBLES = COOP_MAXBP * 8;
  static constexpr int P_DOUBLES = NA * NB;
  static constexpr int JC = (NA * NB < 18) ? NA * NB : 18;  // J_ab elements reduced across the warps per round
  static constexpr int J_DOUBLES = NW * JC * 32;
  // + 2 rows of the density-bound table as int16 (up to DL_SMEM_MAX_NSH shells), see the kernel
  static constexpr size_t SMEM_BYTES = sizeof(double) * (R_DOUBLES + E_DOUBLES + B_DOUBLES + P_DOUBLES + J_DOUBLES) + 2 * DL_SMEM_MAX_NSH * sizeof(short);
};

// ---- R_tuv columns.  S1[n][v] = R^n_{0,0,v}; column (u,v): R^0_{t,u,v}, t = 0..L-u-v, built from the v-th
// two-index triangle A[n][u] = R^n_{0,u,v}.  Template recursion (not "#pragma unroll" on five nested triangular loops,
// which nvcc gives up on and demotes the arrays to local memory) keeps every index a compile-time constant; a warp only
// executes the columns it owns (warp-uniform branches on w).
template <int L, int U, int V>
__device__ __forceinline__ void r_column(const double (&S1)[L + 1][L + 1], double X, double Y, double *__restrict__ Rl) {
  constexpr int M = L - U - V;
  // two-index triangle of this v up to column U: A[n][uu], n + uu <= L - V
  double A[L + 1][U + 1];
#pragma unroll
  for (int n = 0; n <= L - V; ++n) A[n][0] = S1[n][V];
#pragma unroll
  for (int uu = 1; uu <= U; ++uu) {
#pragma unroll
    for (int n = 0; n <= L - V - uu; ++n) {
      double val = Y * A[n + 1][uu - 1];
      if (uu > 1) val = fma((double)(uu - 1), A[n + 1][uu > 1 ? uu - 2 : 0], val);
      A[n][uu] = val;
    }
  }
  // t recursion on the column: B[n][t], n + t <= M
  double B[M + 1][M + 1];
#pragma unroll
  for (int n = 0; n <= M; ++n) B[n][0] = A[n][U];
#pragma unroll
  for (int t = 1; t <= M; ++t) {
#pragma unroll
    for (int n = 0; n <= M - t; ++n) {
      double val = X * B[n + 1][t - 1];
      if (t > 1) val = fma((double)(t - 1), B[n + 1][t > 1 ? t - 2 : 0], val);
      B[n][t] = val;
    }
  }
#pragma unroll
  for (int t = 0; t <= M; ++t) Rl[hidx(L, t, U, V) * 32] = B[0][t];
}

// columns in order of decreasing length m = L-u-v (so that dealing them round-robin balances the warps):
// column K -> (m, v): the m-th group has L-m+1 columns
__host__ __device__ constexpr int col_m(int L, int K) {
  int m = L, k = K;
  while (k >= L - m + 1) { k -= L - m + 1; --m; }
  return m;
}
__host__ __device__ constexpr int col_v(int L, int K) {
  int m = L, k = K;
  while (k >= L - m + 1) { k -= L - m + 1; --m; }
  return k;
}

template <int L, int NW, int K>
struct ColumnDeal {
  static __device__ __forceinline__ void run(const double (&S1)[L + 1][L + 1], double X, double Y, int w, double *__restrict__ Rl) {
    constexpr int M = col_m(L, K), V = col_v(L, K), U = L - M - V;
    if ((K % NW) == w) r_column<L, U, V>(S1, X, Y, Rl);
    if constexpr (K + 1 < (L + 1) * (L + 2) / 2) ColumnDeal<L, NW, K + 1>::run(S1, X, Y, w, Rl);
  }
};

template <int L, int NW>
__device__ __forceinline__ void build_R_columns(const double *Fs, double X, double Y, double Z, int w, double *__restrict__ Rl) {
  double S1[L + 1][L + 1];
#pragma unroll
  for (int n = 0; n <= L; ++n) S1[n][0] = Fs[n];
#pragma unroll
  for (int v = 1; v <= L; ++v) {
#pragma unroll
    for (int n = 0; n <= L - v; ++n) {
      double val = Z * S1[n + 1][v - 1];
      if (v > 1) val = fma((double)(v - 1), S1[n + 1][v > 1 ? v - 2 : 0], val);
      S1[n][v] = val;
    }
  }
  ColumnDeal<L, NW, 0>::run(S1, X, Y, w, Rl);
}

__host__ __device__ constexpr bool r_has_target(int r, int s, int w, int bx, int by, int bz, int LAB) {
  for (int tt = 0; tt <= bx; ++tt)
    for (int uu = 0; uu <= by; ++uu)
      for (int vv = 0; vv <= bz; ++vv) {
        const int t = r - tt, u = s - uu, v = w - vv;
        if (t >= 0 && u >= 0 && v >= 0 && t + u + v <= LAB) return true;
      }
  return false;
}

// ---- one thread's share of a primitive quartet, ket half: G[t u v] = sum E^{c_IC d_ID}_{t'u'v'} R_{t+t',u+u',v+v'}.  This is
// the only code that differs between the warps of a CTA (one fully unrolled body per ket-c component)
template <int LA, int LB, int LC, int LD, int IC, int ID>
__device__ __forceinline__ void ket_contract_cd(const double *__restrict__ Rl, double QCx, double QDx, double QCy, double QDy,
                                                double QCz, double QDz, double oo2q, double (&G)[nherm(LA + LB)]) {
  using S = CoopShape<LA, LB, LC, LD>;
  constexpr int L = S::L, LAB = S::LAB;
  // ket Hermite coefficients (sign folded): only the rows of this thread's component pair survive dead-code elimination
  double Ecd[3][LC + 1][LD + 1][LC + LD + 1];
  unr::hermite_E<LC, LD, true>(Ecd[0], QCx, QDx, oo2q);
  unr::hermite_E<LC, LD, true>(Ecd[1], QCy, QDy, oo2q);
  unr::hermite_E<LC, LD, true>(Ecd[2], QCz, QDz, oo2q);
  constexpr int cx = cart_x(LC, IC), cy = cart_y(LC, IC), cz = LC - cx - cy;
  constexpr int dx = cart_x(LD, ID), dy = cart_y(LD, ID), dz = LD - dx - dy;
  constexpr int bx = cx + dx, by = cy + dy, bz = cz + dz;
  double ek[bx + 1][by + 1][bz + 1];
#pragma unroll
  for (int tt = 0; tt <= bx; ++tt)
#pragma unroll
    for (int uu = 0; uu <= by; ++uu) {
      const double exy = Ecd[0][cx][dx][tt] * Ecd[1][cy][dy][uu];
#pragma unroll
      for (int vv = 0; vv <= bz; ++vv) ek[tt][uu][vv] = exy * Ecd[2][cz][dz][vv];
    }
  // ket contraction, scatter form: every needed R element is read from shared memory once
#pragma unroll
  for (int i = 0; i < nherm(LAB); ++i) G[i] = 0.0;
  static_for<nherm(L)>([&](auto ric) __attribute__((always_inline)) {
    constexpr int ridx = decltype(ric)::value;
    constexpr int r = herm_t(L, ridx), s = herm_u(L, ridx), w = herm_v(L, ridx);
    if constexpr (r_has_target(r, s, w, bx, by, bz, LAB)) {
      const double rv = Rl[ridx * 32];
#pragma unroll
      for (int tt = 0; tt <= bx; ++tt)
#pragma unroll
        for (int uu = 0; uu <= by; ++uu)
#pragma unroll
          for (int vv = 0; vv <= bz; ++vv) {
            const int t = r - tt, u = s - uu, v = w - vv;
            if (t >= 0 && u >= 0 && v >= 0 && t + u + v <= LAB)
              G[hidx(LAB, t, u, v)] = fma(ek[tt][uu][vv], rv, G[hidx(LAB, t, u, v)]);
          }
    }
  });
}

// ---- bra half, the SAME code for every warp: acc[a + NA*b] += sum_tuv E^{ab}_{tuv} G[t u v]; E^ab is CTA-uniform (shared
// memory, broadcast loads).  Keeping it out of the per-component bodies halves the kernel's code (ncu: the cooperative
// kernels stalled on instruction fetch, "no_instruction" 5-7 per issue, with six different bodies in flight per CTA)
template <int LA, int LB>
__device__ __forceinline__ void bra_contract(const double *__restrict__ Es, const double (&G)[nherm(LA + LB)], double *acc) {
  constexpr int LAB = LA + LB, NA = ncart(LA), NE1 = (LA + 1) * (LB + 1) * (LAB + 1);
  const double *__restrict__ Ex = Es, *__restrict__ Ey = Es + NE1, *__restrict__ Ez = Es + 2 * NE1;
  static_for<NA * ncart(LB)>([&](auto abc) __attribute__((always_inline)) {
    constexpr int iab = decltype(abc)::value, ia = iab % NA, ib = iab / NA;
    constexpr int ax = cart_x(LA, ia), ay = cart_y(LA, ia), az = LA - ax - ay;
    constexpr int bbx = cart_x(LB, ib), bby = cart_y(LB, ib), bbz = LB - bbx - bby;
    // straight into the accumulator (no s = 0; ...; acc += s), and the E_000 = 1 factors of an axis without angular
    // momentum are skipped at compile time: the staged coefficients are run-time values, so nvcc cannot fold them
    constexpr bool one_x = (ax + bbx == 0), one_y = (ay + bby == 0), one_z = (az + bbz == 0);
    double s = acc[iab];
#pragma unroll
    for (int t = 0; t <= ax + bbx; ++t)
#pragma unroll
      for (int u = 0; u <= ay + bby; ++u) {
        double sv;
        if constexpr (one_z) sv = G[hidx(LAB, t, u, 0)];
        else {
          sv = Ez[(az * (LB + 1) + bbz) * (LAB + 1)] * G[hidx(LAB, t, u, 0)];
#pragma unroll
          for (int v = 1; v <= az + bbz; ++v) sv = fma(Ez[(az * (LB + 1) + bbz) * (LAB + 1) + v], G[hidx(LAB, t, u, v)], sv);
        }
        if constexpr (one_x && one_y) s += sv;
        else if constexpr (one_x) s = fma(Ey[(ay * (LB + 1) + bby) * (LAB + 1) + u], sv, s);
        else if constexpr (one_y) s = fma(Ex[(ax * (LB + 1) + bbx) * (LAB + 1) + t], sv, s);
        else s = fma(Ex[(ax * (LB + 1) + bbx) * (LAB + 1) + t] * Ey[(ay * (LB + 1) + bby) * (LAB + 1) + u], sv, s);
      }
    acc[iab] = s;
  });
}

// ---- J/K digestion of one thread's (c_IC, d_ID) slice: the ket-indexed updates leave as reductions, the J_ab
// contribution I*P_cd is ADDED to tab[] (reduced over warps and lanes by the caller)
// non-axial factor of Cartesian component `idx` (runtime) of a shell with angular momentum L (compile time)
template <int L>
__device__ __forceinline__ double nonaxial_idx(int idx) {
  if (L < 2) return 1.0;
  if (L == 2) return (idx == 1 || idx == 2 || idx == 4) ? 1.7320508075688772 : 1.0;  // xy, xz, yz
  int i = 0, x = L, y = 0;
  for (int xx = L; xx >= 0; --xx)
    for (int yy = L - xx; yy >= 0; --yy, ++i)
      if (i == idx) { x = xx; y = yy; }
  return nonaxial(L, x, y, L - x - y);
}

// (the ket-c component ic is a RUNTIME value here: one copy of this code serves every warp)
template <int LA, int LB, int LC, int LD, int ID>
__device__ __forceinline__ void digest_cd(const double *acc, bool active, double deg, int fa, int fb, int fc, int fd, int ic,
                                          const double *__restrict__ Pab, const ClassParams &prm, double *tab) {
  constexpr int NA = ncart(LA), NB = ncart(LB);
  constexpr int dx = cart_x(LD, ID), dy = cart_y(LD, ID);
  if (!active) return;  // (warp-divergent; no collective below)
  const int N = prm.nbf;
  const double *__restrict__ P = prm.P;
  const int gc = fc + ic, gd = fd + ID;
  const double ncd = nonaxial_idx<LC>(ic) * nonaxial(LD, dx, dy, LD - dx - dy) * deg;
  const double pcd = __ldg(P + gc + (size_t)N * gd);
  double jcd = 0.0;
  double Kac[NA], Kad[NA], Kbc[NB], Kbd[NB];
#pragma unroll
  for (int i = 0; i < NA; ++i) Kac[i] = Kad[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NB; ++i) Kbc[i] = Kbd[i] = 0.0;
  double pad[NA], pac[NA];
#pragma unroll
  for (int ia = 0; ia < NA; ++ia) {
    pad[ia] = __ldg(P + gd + (size_t)N * (fa + ia));
    pac[ia] = __ldg(P + gc + (size_t)N * (fa + ia));
  }
  static_for<NB>([&](auto ibc) __attribute__((always_inline)) {
    constexpr int ib = decltype(ibc)::value;
    constexpr int bx = cart_x(LB, ib), by = cart_y(LB, ib);
    const double nb = nonaxial(LB, bx, by, LB - bx - by) * ncd;
    const double pbd = __ldg(P + gd + (size_t)N * (fb + ib));
    const double pbc = __ldg(P + gc + (size_t)N * (fb + ib));
    static_for<NA>([&](auto iac) __attribute__((always_inline)) {
      constexpr int ia = decltype(iac)::value;
      constexpr int ax = cart_x(LA, ia), ay = cart_y(LA, ia);
      const double v = acc[ia + NA * ib] * (nonaxial(LA, ax, ay, LA - ax - ay) * nb);
      jcd = fma(v, Pab[ia + NA * ib], jcd);
      Kac[ia] = fma(v, pbd, Kac[ia]);
      Kad[ia] = fma(v, pbc, Kad[ia]);
      Kbc[ib] = fma(v, pad[ia], Kbc[ib]);
      Kbd[ib] = fma(v, pac[ia], Kbd[ib]);
      tab[ia + NA * ib] = fma(v, pcd, tab[ia + NA * ib]);
    });
  });
  atomicAdd(prm.J + gc + (size_t)N * gd, 0.5 * jcd);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    atomicAdd(prm.K + gc + (size_t)N * (fa + i), 0.25 * Kac[i]);
    atomicAdd(prm.K + gd + (size_t)N * (fa + i), 0.25 * Kad[i]);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    atomicAdd(prm.K + gc + (size_t)N * (fb + i), 0.25 * Kbc[i]);
    atomicAdd(prm.K + gd + (size_t)N * (fb + i), 0.25 * Kbd[i]);
  }
}

// ---- tensor mode: one thread's (c,d) slice into the N^4 tensor at its 8 symmetric positions
template <int LA, int LB, int LC, int LD, int ID>
__device__ __forceinline__ void scatter_cd(const double *acc, int fa, int fb, int fc, int fd, int ic, const ClassParams &prm) {
  constexpr int NA = ncart(LA), NB = ncart(LB);
  constexpr int dx = cart_x(LD, ID), dy = cart_y(LD, ID);
  const size_t N = prm.nbf;
  double *T = prm.tensor;
  const double ncd = nonaxial_idx<LC>(ic) * nonaxial(LD, dx, dy, LD - dx - dy);
  const size_t c = fc + ic, d = fd + ID;
  static_for<NA * NB>([&](auto abc) __attribute__((always_inline)) {
    constexpr int iab = decltype(abc)::value, ia = iab % NA, ib = iab / NA;
    constexpr int ax = cart_x(LA, ia), ay = cart_y(LA, ia), bx = cart_x(LB, ib), by = cart_y(LB, ib);
    const double I = acc[iab] * (nonaxial(LA, ax, ay, LA - ax - ay) * nonaxial(LB, bx, by, LB - bx - by) * ncd);
    const size_t a = fa + ia, b = fb + ib;
    T[a + N * (b + N * (c + N * d))] = I;
    T[b + N * (a + N * (c + N * d))] = I;
    T[a + N * (b + N * (d + N * c))] = I;
    T[b + N * (a + N * (d + N * c))] = I;
    T[c + N * (d + N * (a + N * b))] = I;
    T[d + N * (c + N * (a + N * b))] = I;
    T[c + N * (d + N * (b + N * a))] = I;
    T[d + N * (c + N * (b + N * a))] = I;
  });
}

// warp w runs the code of ket-c components w, w+NW, ..: f(integral_constant<IC>, integral_constant<slot>).  Per slot the
// NW candidates form ONE if / else-if chain (mutually exclusive branches: the register allocator must not think that
// several of the fully unrolled component bodies can run back to back).
template <int NW, int SLOT, int W = 0>
struct CoopChain {
  template <class F>
  static __device__ __forceinline__ void run(int w, F &&f) {
    if (w == W) f(std::integral_constant<int, SLOT * NW + W>(), std::integral_constant<int, SLOT>());
    else if constexpr (W + 1 < NW) CoopChain<NW, SLOT, W + 1>::run(w, f);
  }
};
template <int NC, int NW, int SLOT = 0>
struct CoopDispatch {
  template <class F>
  static __device__ __forceinline__ void run(int w, F &&f) {
    CoopChain<NW, SLOT>::run(w, f);
    if constexpr ((SLOT + 1) * NW < NC) CoopDispatch<NC, NW, SLOT + 1>::run(w, f);
  }
};

// prm.tensor == nullptr: J/K digestion with screening; else tensor scatter (no screening).
// DSEL: the ket-d component of this launch; CPT: ket-c components per thread.
template <int LA, int LB, int LC, int LD, int DSEL, int CPT, int MINB>
__global__ void __launch_bounds__(32 * (ncart(LC) / CPT), MINB) eri_coop_kernel(const __grid_constant__ ClassParams prm) {
  using S = CoopShape<LA, LB, LC, LD, CPT>;
  constexpr int NW = S::NW, NC = ncart(LC), L = S::L, NA = S::NA, NB = S::NB, NAB = NA * NB, NE1 = S::NE1, JC = S::JC;
  constexpr int QUEUE_NBK = 4;
  __shared__ int s_queue[NW][QUEUE_NBK * 64];  // per-warp (identical) queues of surviving ket pairs, bucketed by contraction
  extern __shared__ double coop_smem[];
  double *Rs = coop_smem;
  double *Es = Rs + S::R_DOUBLES;
  double *Bs = Es + S::E_DOUBLES;
  double *Pab = Bs + S::B_DOUBLES;
  double *Js = Pab + S::P_DOUBLES;
  const int lane = threadIdx.x, w = threadIdx.y, tid = w * 32 + lane;
  const bool tensor_mode = prm.tensor != nullptr;
  const double *s_boys = prm.boys + (size_t)L * BOYS_TABLE_DOUBLES;
  const int *__restrict__ dl = prm.dl;
  const int nsh = prm.nsh;
  const int dlmax = prm.screen ? dl[nsh * nsh] : 0;  // device-resident density bound (see eri_body.inc)
  unsigned n_q = 0, n_pq = 0;
  double *Rl = Rs + lane;
  // CTA = one bra pair x every ysplit-th ket tile (32 ket pairs each); bra pairs are walked with a grid stride
  const int y = blockIdx.x % prm.ysplit;
  const int row_step = max(1, (int)gridDim.x / prm.ysplit);
#pragma unroll 1
  for (int rowl = blockIdx.x / prm.ysplit;; rowl += row_step) {
  const int ib = rowl * prm.nranks + prm.rank;
  if (ib >= prm.bra_n) break;
  {
    const int ik0 = y * TILE_KET;
    if (ik0 >= prm.ket_n) break;
    if (prm.same && ik0 > ib) continue;
    if (prm.screen && prm.pair_qls[prm.bra_off + ib] + prm.pair_qls[prm.ket_off + ik0] + dlmax < prm.tlog) break;
  }
  const int gb = prm.bra_off + ib;
  const PairInfo bi = load_pair_info(prm.pair_info, gb);
  const int sa = bi.a, sb = bi.b;
  const int qlb = bi.ql;
  const int bra_off = bi.poff, bra_np = bi.np;
  const double Ax = bi.Ax, Ay = bi.Ay, Az = bi.Az;
  const double Bx = bi.Bx, By = bi.By, Bz = bi.Bz;
  const int fa = bi.fa, fb = bi.fb;
  const double degab = (sa == sb) ? 1.0 : 2.0;
  const int dab = prm.screen ? dl[sa + nsh * sb] : 0;
  __syncthreads();  // the previous bra pair's readers of Pab / E^ab are done
  if (!tensor_mode)
    for (int i = tid; i < NAB; i += 32 * NW) Pab[i] = prm.P[(fa + i % NA) + (size_t)prm.nbf * (fb + i / NA)];
  // rows dl[., sa], dl[., sb] of the density-bound table for the screening walk (see eri_body.inc)
  const bool dl_sm = prm.screen && prm.dl_smem;
  short *dla = reinterpret_cast<short *>(Js + S::J_DOUBLES), *dlb = dla + nsh;
  if (dl_sm) {
    for (int i = tid; i < nsh; i += 32 * NW) {
      dla[i] = (short)max(dl[i + nsh * sa], -32768);
      dlb[i] = (short)max(dl[i + nsh * sb], -32768);
    }
    __syncthreads();
  }
  int chunk_loaded = -1;

  // ---- one batch of (up to 32) quartets of the CTA's bra pair: lane -> ket pair ik of the class (identical in every warp)
  // Survivor compaction as in the register flavour (eri_body.inc): every warp of the CTA runs the same screening walk and
  // keeps its own, identical copy of the queue, so all warps call process() the same number of times (it synchronises).
  unsigned qn = 0;  // queued survivors per contraction bucket, 8 bits each (identical in every warp)
  int *queue = s_queue[w];
  // one loop, one textual copy of the batch body: the last trip only drains the queues
#pragma unroll 1
  for (int bxk = y;; bxk += prm.ysplit) {
    const int ik0 = bxk * TILE_KET;
    bool last = bxk >= prm.nbx || (prm.same && ik0 > ib);
    if (!last && prm.screen && prm.pair_qls[gb] + prm.pair_qls[prm.ket_off + ik0] + dlmax < prm.tlog) last = true;
    const int ik = ik0 + lane;
    bool active = false;
    int bucket = 0;
    if (!last) {
      active = ik < prm.ket_n && (!prm.same || ik <= ib);
      if (prm.screen && active) {
        const int gk = prm.ket_off + ik;
        const int s = qlb + prm.pair_ql[gk];
        if (s + dlmax < prm.tlog) active = false;
        else {
          const int sc = prm.pair_a[gk], sd = prm.pair_b[gk];
          int dm = max(dab, prm.pair_dl[gk]);
          if (dl_sm) {
            dm = max(dm, max((int)dla[sc], (int)dla[sd]));
            dm = max(dm, max((int)dlb[sc], (int)dlb[sd]));
          } else {
            dm = max(dm, max(dl[sc + nsh * sa], dl[sd + nsh * sa]));
            dm = max(dm, max(dl[sc + nsh * sb], dl[sd + nsh * sb]));
          }
          active = (s + dm >= prm.tlog);
        }
      }
      if (!tensor_mode && active) {
        const int np = prm.pair_np[prm.ket_off + ik];
        bucket = (np > 1) + (np > 4) + (np > 12);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, active);  // identical in every warp of the CTA
    if (!tensor_mode) {
#pragma unroll
      for (int k = 0; k < QUEUE_NBK; ++k) {
        const unsigned mk = __ballot_sync(0xffffffffu, active && bucket == k);
        const int cnt = (qn >> (8 * k)) & 0xffu;
        if (active && bucket == k) queue[64 * k + cnt + __popc(mk & ((1u << lane) - 1u))] = ik;
        qn += (unsigned)__popc(mk) << (8 * k);
      }
      __syncwarp();
    }
    int mcur = 0;  // cursor of the merged drain at the end of the walk
#pragma unroll 1
    for (bool first = true;; first = false) {  // drain: every bucket that holds a full batch; at the end of the walk the rest
    int ikc = ik;
    bool act = active, run = first && m != 0u;
    if (!tensor_mode) {
      int kb = -1;
#pragma unroll
      for (int k = QUEUE_NBK - 1; k >= 0; --k) {
        const int cnt = (qn >> (8 * k)) & 0xffu;
        if (cnt >= 32) kb = k;
      }
      run = kb >= 0;
      if (run) {
        const int cnt = (qn >> (8 * kb)) & 0xffu;
        const int rest = cnt - 32;
        int *qb = queue + 64 * kb;
        act = true;
        ikc = qb[lane];
        const int carry = (lane < rest) ? qb[32 + lane] : 0;
        __syncwarp();
        if (lane < rest) qb[lane] = carry;
        qn -= 32u << (8 * kb);
        __syncwarp();
      } else if (last) {
        // end of the walk: the leftovers of all buckets, longest contraction first, in batches of 32 (see eri_body.inc)
        const int c3 = (qn >> 24) & 0xffu, c2 = (qn >> 16) & 0xffu, c1 = (qn >> 8) & 0xffu, c0 = qn & 0xffu;
        const int o3 = c3, o2 = o3 + c2, o1 = o2 + c1, total = o1 + c0;
        run = mcur < total;
        if (run) {
          const int e = mcur + lane;
          act = e < total;
          const int src = e < o3 ? 192 + e : (e < o2 ? 128 + (e - o3) : (e < o1 ? 64 + (e - o2) : e - o1));
          ikc = act ? queue[src] : 0;
          mcur += 32;
        }
      }
    }
    if (!run) break;
    {  // one batch: lane -> ket pair ikc of the class (textually inlined: nvcc does not inline a lambda this large)
      const int ik = ikc;
      const bool active = act;
      const int gk = prm.ket_off + min(max(ik, 0), prm.ket_n - 1);
      const PairInfo ki = load_pair_info(prm.pair_info, gk);
      const int sc = ki.a, sd = ki.b;
      const int ket_off = ki.poff;
      const int ket_np = ki.np;
      const int kmax = __reduce_max_sync(0xffffffffu, active ? ket_np : 0);
      if (active && w == 0) { n_q += 1u; n_pq += (unsigned)(bra_np * ket_np); }
      const double Cx = ki.Ax, Cy = ki.Ay, Cz = ki.Az;
      const double Dx = ki.Bx, Dy = ki.By, Dz = ki.Bz;
      double acc[CPT][NAB];
  #pragma unroll
      for (int k = 0; k < CPT; ++k)
  #pragma unroll
        for (int i = 0; i < NAB; ++i) acc[k][i] = 0.0;

  #pragma unroll 1
      for (int ip0 = 0; ip0 < bra_np; ip0 += COOP_MAXBP) {
        const int nchunk = min(COOP_MAXBP, bra_np - ip0);
        if (chunk_loaded != ip0) {  // CTA-uniform
          __syncthreads();
          if (tid < 3 * nchunk) {
            const int pi = tid / 3, axis = tid - 3 * pi;
            const double2 *rb = prm.pp.rec + (size_t)bra_off + (size_t)(3 * TILE_KET) * (ip0 + pi);
            const double2 b0 = __ldg(rb), b1 = __ldg(rb + TILE_KET), b2 = __ldg(rb + 2 * TILE_KET);
            const double Pc = axis == 0 ? b0.y : (axis == 1 ? b1.x : b1.y);
            const double Ac = axis == 0 ? Ax : (axis == 1 ? Ay : Az);
            const double Bc = axis == 0 ? Bx : (axis == 1 ? By : Bz);
            double E[LA + 1][LB + 1][LA + LB + 1];
            unr::hermite_E<LA, LB, false>(E, Pc - Ac, Pc - Bc, b2.y);
            double *dst = Es + (size_t)(pi * 3 + axis) * NE1;
  #pragma unroll
            for (int i = 0; i <= LA; ++i)
  #pragma unroll
              for (int j = 0; j <= LB; ++j)
  #pragma unroll
                for (int t = 0; t <= LA + LB; ++t) dst[(i * (LB + 1) + j) * (LA + LB + 1) + t] = (t <= i + j) ? E[i][j][t] : 0.0;
            if (axis == 0) {
              double *br = Bs + pi * 8;
              br[0] = b0.x; br[1] = b0.y; br[2] = b1.x; br[3] = b1.y; br[4] = b2.x;
            }
          }
          chunk_loaded = ip0;
          __syncthreads();
        }
  #pragma unroll 1
        for (int pi = 0; pi < nchunk; ++pi) {
          const double p = Bs[pi * 8], Px = Bs[pi * 8 + 1], Py = Bs[pi * 8 + 2], Pz = Bs[pi * 8 + 3], kab = Bs[pi * 8 + 4];
          const double *Ep = Es + (size_t)(pi * 3) * NE1;
  #pragma unroll 1
          for (int iq = 0; iq < kmax; ++iq) {
            const bool on = active && iq < ket_np;
            const double2 *rk = prm.pp.rec + (size_t)ket_off + (size_t)(3 * TILE_KET) * min(iq, ket_np - 1);
            const double2 k0 = __ldg(rk), k1 = __ldg(rk + TILE_KET), k2 = __ldg(rk + 2 * TILE_KET);
            const double q = k0.x, Qx = k0.y, Qy = k1.x, Qz = k1.y;
            const double r = rsqrt(p + q);
            const double alpha = p * q * (r * r);
            const double X = Px - Qx, Y = Py - Qy, Z = Pz - Qz;
            const double T = alpha * (X * X + Y * Y + Z * Z);
            const double pref = on ? QLB_TWO_PI_52 * kab * k2.x * r : 0.0;
            double F[L + 1];
            boys_array<L>(T, s_boys, F);
            {
              double sgn = pref;  // pref * (-2 alpha)^n
  #pragma unroll
              for (int n = 0; n <= L; ++n) {
                F[n] *= sgn;
                sgn *= -2.0 * alpha;
              }
            }
            build_R_columns<L, NW>(F, X, Y, Z, w, Rl);
            __syncthreads();
            static_for<CPT>([&](auto slotc) __attribute__((always_inline)) {
              constexpr int slot = decltype(slotc)::value;
              double G[nherm(LA + LB)];
              CoopChain<NW, slot>::run(w, [&](auto icc, auto) __attribute__((always_inline)) {
                ket_contract_cd<LA, LB, LC, LD, decltype(icc)::value, DSEL>(Rl, Qx - Cx, Qx - Dx, Qy - Cy, Qy - Dy, Qz - Cz, Qz - Dz, k2.y, G);
              });
              bra_contract<LA, LB>(Ep, G, acc[slot]);
            });
            __syncthreads();
          }
        }
      }
      // ---- digestion / scatter of this thread's slices
      const int fc = ki.fa, fd = ki.fb;
      if (!tensor_mode) {
        const double deg = degab * ((sc == sd) ? 1.0 : 2.0) * ((prm.same && ib == ik) ? 1.0 : 2.0);
        double tab[NAB];  // this thread's J_ab contribution (all its ket-c components)
  #pragma unroll
        for (int i = 0; i < NAB; ++i) tab[i] = 0.0;
#pragma unroll
        for (int slot = 0; slot < CPT; ++slot)
          digest_cd<LA, LB, LC, LD, DSEL>(acc[slot], active, deg, fa, fb, fc, fd, slot * NW + w, Pab, prm, tab);
        // J_ab: sum over the warps through shared memory, then over the lanes; JC elements per round
        static_for<(NAB + JC - 1) / JC>([&](auto rc) __attribute__((always_inline)) {
          constexpr int c0 = decltype(rc)::value * JC;
          static_for<JC>([&](auto ic) __attribute__((always_inline)) {
            constexpr int i = decltype(ic)::value;
            Js[(w * JC + i) * 32 + lane] = (c0 + i < NAB) ? tab[c0 + i < NAB ? c0 + i : 0] : 0.0;
          });
          __syncthreads();
          for (int i = w; i < JC && c0 + i < NAB; i += NW) {
            double s = 0.0;
  #pragma unroll
            for (int ww = 0; ww < NW; ++ww) s += Js[(ww * JC + i) * 32 + lane];
            s = warp_sum(s);
            const int iab = c0 + i;
            if (lane == 0 && s != 0.0) atomicAdd(prm.J + (fa + iab % NA) + (size_t)prm.nbf * (fb + iab / NA), 0.5 * s);
          }
          __syncthreads();
        });
      } else if (active) {
#pragma unroll
        for (int slot = 0; slot < CPT; ++slot) scatter_cd<LA, LB, LC, LD, DSEL>(acc[slot], fa, fb, fc, fd, slot * NW + w, prm);
      }
    }
    if (tensor_mode) break;  // no queues: one direct batch per tile
    }  // drain
    if (last) break;
  }
  }  // bra pairs
  if (prm.counters && DSEL == 0 && w == 0) {
    n_q = __reduce_add_sync(0xffffffffu, n_q);
    n_pq = __reduce_add_sync(0xffffffffu, n_pq);
    if (lane == 0 && n_q) {
      atomicAdd(prm.counters + 0, (unsigned long long)n_q);
      atomicAdd(prm.counters + 1, (unsigned long long)n_pq);
    }
  }
}

}  // namespace cop
}  // namespace qlb
