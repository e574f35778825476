// md_core.cuh -- device-side McMurchie-Davidson machinery shared by every ERI class kernel.
//
// Replaces, per primitive quartet, the reference's closed-form evaluation
//   computeElectronRepulsionIntegral (IntegralsModule.jl:412-445) + BFactors/thetaFactors (:354-409) +
//   FIntegral (:283-289, Boys via GSL) + GaussProductPolynomialFactor (:91-120)
// with the algebraically identical (checked to 1e-16 by the oracle tests) Hermite expansion
//   (ab|cd) = 2 pi^2.5 / (p q sqrt(p+q)) * sum_tuv E^ab_tuv sum_t'u'v' (-1)^(t'+u'+v') E^cd_t'u'v' R_(t+t',u+u',v+v').
// Everything is templated on the four shell angular momenta so that, with UNR=true, every loop unrolls and all
// intermediates (R, E, G, accumulators) live in registers; with UNR=false the same source runs as rolled loops
// over local-memory arrays (used for the high-L classes whose straight-line code would not fit the I-cache).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb {

#define QLB_PI 3.14159265358979323846
#define QLB_TWO_PI_52 34.98683665524972497  // 2*pi^2.5

__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int nherm(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }
// position of (t,u,v), t+u+v <= L, in the tetrahedral layout of total order L
__host__ __device__ constexpr int hidx(int L, int t, int u, int v) {
  return nherm(L) - nherm(L - t) + u * (L - t + 1) - u * (u - 1) / 2 + v;
}

// ------------------------------------------------------------------ Boys function
// Table: rows T_i = i*BOYS_DT (i = 0..BOYS_NROW-1), columns m = 0..BOYS_NCOL-1 holding F_m(T_i).
// F_L(T) by an 8-term Taylor expansion about the nearest row, lower orders by downward recursion;
// beyond the table by the asymptotic F_0 and upward recursion (exp(-T) kept).  Accurate to ~1e-15 relative.
constexpr int BOYS_NCOL = 24;           // supports total L <= 16
constexpr double BOYS_DT = 0.1;
constexpr double BOYS_TMAX = 36.0;
constexpr int BOYS_NROW = 362;          // rows 0..361 (T up to 36.1)

template <int L>
__device__ __forceinline__ void boys_array(double T, const double *__restrict__ tab, double *F) {
  const double e = exp(-T);
  if (T < BOYS_TMAX) {
    const int i = (int)(T * (1.0 / BOYS_DT) + 0.5);
    const double d = (double)i * BOYS_DT - T;  // T_i - T
    const double *row = tab + i * BOYS_NCOL + L;
    double f = __ldg(row + 7);
#pragma unroll
    for (int k = 6; k >= 0; --k) f = fma(f, d * (1.0 / (k + 1)), __ldg(row + k));
    F[L] = f;
    const double t2 = 2.0 * T;
#pragma unroll
    for (int m = L; m >= 1; --m) F[m - 1] = fma(t2, F[m], e) * (1.0 / (2 * m - 1));
  } else {
    const double it = 1.0 / T;
    F[0] = 0.5 * sqrt(QLB_PI * it);
    const double i2t = 0.5 * it;
#pragma unroll
    for (int m = 0; m < L; ++m) F[m + 1] = fma((double)(2 * m + 1), F[m], -e) * i2t;
  }
}

// ------------------------------------------------------------------ Hermite expansion coefficients, one axis
// E[i][j][t], i<=LA, j<=LB, t<=i+j  (zero beyond).  SIGN=true folds (-1)^t in (ket side).
template <int LA, int LB, bool UNR, bool SIGN>
__device__ __forceinline__ void hermite_E(double (&E)[LA + 1][LB + 1][LA + LB + 1], double PA, double PB, double oo2p) {
  E[0][0][0] = 1.0;
#pragma unroll(UNR ? 8 : 1)
  for (int t = 1; t <= LA + LB; ++t) E[0][0][t] = 0.0;
#pragma unroll(UNR ? 8 : 1)
  for (int i = 0; i <= LA; ++i) {
#pragma unroll(UNR ? 8 : 1)
    for (int j = 0; j <= LB; ++j) {
      if (i == 0 && j == 0) continue;
#pragma unroll(UNR ? 8 : 1)
      for (int t = 0; t <= LA + LB; ++t) {
        double v = 0.0;
        if (t <= i + j) {
          if (j == 0) {  // raise i from (i-1, 0)
            v = PA * E[i - 1][0][t];
            if (t > 0) v = fma(oo2p, E[i - 1][0][t - 1], v);
            if (t + 1 <= i - 1) v = fma((double)(t + 1), E[i - 1][0][t + 1], v);
          } else {  // raise j from (i, j-1)
            v = PB * E[i][j - 1][t];
            if (t > 0) v = fma(oo2p, E[i][j - 1][t - 1], v);
            if (t + 1 <= i + j - 1) v = fma((double)(t + 1), E[i][j - 1][t + 1], v);
          }
        }
        E[i][j][t] = v;
      }
    }
  }
  if (SIGN) {
#pragma unroll(UNR ? 8 : 1)
    for (int i = 0; i <= LA; ++i)
#pragma unroll(UNR ? 8 : 1)
      for (int j = 0; j <= LB; ++j)
#pragma unroll(UNR ? 8 : 1)
        for (int t = 1; t <= LA + LB; t += 2) E[i][j][t] = -E[i][j][t];
  }
}

// ------------------------------------------------------------------ Hermite Coulomb integrals R_tuv, t+u+v <= L
// In: F[n] = pref * F_n(T), n = 0..L; alpha; PQ.  Out: R[hidx(L,t,u,v)] = pref * R^0_tuv.
// In place: level n (orders <= L-n) overwrites level n+1 from the highest order down.
template <int L, bool UNR>
__device__ __forceinline__ void hermite_R(const double *F, double alpha, double X, double Y, double Z, double *R) {
  double s = 1.0;  // (-2 alpha)^n
  double Fs[L + 1];
#pragma unroll
  for (int n = 0; n <= L; ++n) {
    Fs[n] = F[n] * s;
    s *= -2.0 * alpha;
  }
  R[0] = Fs[L];
#pragma unroll(UNR ? 32 : 1)
  for (int n = L - 1; n >= 0; --n) {
#pragma unroll(UNR ? 32 : 1)
    for (int k = L - n; k >= 1; --k) {
#pragma unroll(UNR ? 32 : 1)
      for (int t = k; t >= 0; --t) {
#pragma unroll(UNR ? 32 : 1)
        for (int u = k - t; u >= 0; --u) {
          const int v = k - t - u;
          double val;
          if (t > 0) {
            val = X * R[hidx(L, t - 1, u, v)];
            if (t > 1) val = fma((double)(t - 1), R[hidx(L, t - 2, u, v)], val);
          } else if (u > 0) {
            val = Y * R[hidx(L, t, u - 1, v)];
            if (u > 1) val = fma((double)(u - 1), R[hidx(L, t, u - 2, v)], val);
          } else {
            val = Z * R[hidx(L, t, u, v - 1)];
            if (v > 1) val = fma((double)(v - 1), R[hidx(L, t, u, v - 2)], val);
          }
          R[hidx(L, t, u, v)] = val;
        }
      }
    }
    R[0] = Fs[n];
  }
}

// ------------------------------------------------------------------ primitive-pair records (SoA in HBM)
struct PrimPairs {
  const double *__restrict__ p;   // exponent sum
  const double *__restrict__ x;   // Gaussian product centre
  const double *__restrict__ y;
  const double *__restrict__ z;
  const double *__restrict__ k;   // c_a c_b exp(-a b/p |AB|^2)
};

struct PairRef {  // what one thread needs to know about one shell pair
  double Ax, Ay, Az, Bx, By, Bz;
  int off, n;     // primitive-pair range
};

// ------------------------------------------------------------------ contracted quartet block accumulation
// acc[a + NA*(b + NB*(c + NC*(d-DLO)))] += (ab|cd) (un-normalised w.r.t. non-axial factors) for d in [DLO,DHI).
template <int LA, int LB, int LC, int LD, bool UNR, int DLO, int DHI>
__device__ __forceinline__ void quartet_accumulate(const PrimPairs &pp, const PairRef &bra, const PairRef &ket,
                                                   const double *__restrict__ boys_tab, double *acc) {
  constexpr int LAB = LA + LB, LCD = LC + LD, L = LAB + LCD;
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC);
#pragma unroll 1
  for (int ip = 0; ip < bra.n; ++ip) {
    const double p = pp.p[bra.off + ip], Px = pp.x[bra.off + ip], Py = pp.y[bra.off + ip], Pz = pp.z[bra.off + ip];
    const double kab = pp.k[bra.off + ip];
    double Eab[3][LA + 1][LB + 1][LAB + 1];
    {
      const double oo2p = 0.5 / p;
      hermite_E<LA, LB, UNR, false>(Eab[0], Px - bra.Ax, Px - bra.Bx, oo2p);
      hermite_E<LA, LB, UNR, false>(Eab[1], Py - bra.Ay, Py - bra.By, oo2p);
      hermite_E<LA, LB, UNR, false>(Eab[2], Pz - bra.Az, Pz - bra.Bz, oo2p);
    }
#pragma unroll 1
    for (int iq = 0; iq < ket.n; ++iq) {
      const double q = pp.p[ket.off + iq], Qx = pp.x[ket.off + iq], Qy = pp.y[ket.off + iq], Qz = pp.z[ket.off + iq];
      const double kcd = pp.k[ket.off + iq];
      double Ecd[3][LC + 1][LD + 1][LCD + 1];
      {
        const double oo2q = 0.5 / q;
        hermite_E<LC, LD, UNR, true>(Ecd[0], Qx - ket.Ax, Qx - ket.Bx, oo2q);
        hermite_E<LC, LD, UNR, true>(Ecd[1], Qy - ket.Ay, Qy - ket.By, oo2q);
        hermite_E<LC, LD, UNR, true>(Ecd[2], Qz - ket.Az, Qz - ket.Bz, oo2q);
      }
      const double pq = p + q;
      const double ipq = 1.0 / pq;
      const double alpha = p * q * ipq;
      const double X = Px - Qx, Y = Py - Qy, Z = Pz - Qz;
      const double T = alpha * (X * X + Y * Y + Z * Z);
      const double pref = QLB_TWO_PI_52 * kab * kcd * sqrt(ipq) / (p * q);
      double F[L + 1];
      boys_array<L>(T, boys_tab, F);
#pragma unroll
      for (int n = 0; n <= L; ++n) F[n] *= pref;
      double R[nherm(L)];
      hermite_R<L, UNR>(F, alpha, X, Y, Z, R);

      // ---- ket contraction then bra contraction, one ket component pair at a time
      int id = 0;
#pragma unroll(UNR ? 16 : 1)
      for (int dx = LD; dx >= 0; --dx) {
#pragma unroll(UNR ? 16 : 1)
        for (int dy = LD - dx; dy >= 0; --dy, ++id) {
          const int dz = LD - dx - dy;
          if (id < DLO || id >= DHI) continue;
          int ic = 0;
#pragma unroll(UNR ? 16 : 1)
          for (int cx = LC; cx >= 0; --cx) {
#pragma unroll(UNR ? 16 : 1)
            for (int cy = LC - cx; cy >= 0; --cy, ++ic) {
              const int cz = LC - cx - cy;
              // products of the three ket E factors for this component pair
              double ek[(LCD + 1) * (LCD + 1) * (LCD + 1)];
              {
                int kk = 0;
#pragma unroll(UNR ? 8 : 1)
                for (int tt = 0; tt <= cx + dx; ++tt)
#pragma unroll(UNR ? 8 : 1)
                  for (int uu = 0; uu <= cy + dy; ++uu) {
                    const double exy = Ecd[0][cx][dx][tt] * Ecd[1][cy][dy][uu];
#pragma unroll(UNR ? 8 : 1)
                    for (int vv = 0; vv <= cz + dz; ++vv) ek[kk++] = exy * Ecd[2][cz][dz][vv];
                  }
              }
              double G[nherm(LAB)];
#pragma unroll(UNR ? 16 : 1)
              for (int t = 0; t <= LAB; ++t)
#pragma unroll(UNR ? 16 : 1)
                for (int u = 0; u <= LAB - t; ++u)
#pragma unroll(UNR ? 16 : 1)
                  for (int v = 0; v <= LAB - t - u; ++v) {
                    double g = 0.0;
                    int kk = 0;
#pragma unroll(UNR ? 8 : 1)
                    for (int tt = 0; tt <= cx + dx; ++tt)
#pragma unroll(UNR ? 8 : 1)
                      for (int uu = 0; uu <= cy + dy; ++uu)
#pragma unroll(UNR ? 8 : 1)
                        for (int vv = 0; vv <= cz + dz; ++vv) g = fma(ek[kk++], R[hidx(L, t + tt, u + uu, v + vv)], g);
                    G[hidx(LAB, t, u, v)] = g;
                  }
              // bra contraction
              int ib = 0;
#pragma unroll(UNR ? 16 : 1)
              for (int bx = LB; bx >= 0; --bx) {
#pragma unroll(UNR ? 16 : 1)
                for (int by = LB - bx; by >= 0; --by, ++ib) {
                  const int bz = LB - bx - by;
                  int ia = 0;
#pragma unroll(UNR ? 16 : 1)
                  for (int ax = LA; ax >= 0; --ax) {
#pragma unroll(UNR ? 16 : 1)
                    for (int ay = LA - ax; ay >= 0; --ay, ++ia) {
                      const int az = LA - ax - ay;
                      double s = 0.0;
#pragma unroll(UNR ? 8 : 1)
                      for (int t = 0; t <= ax + bx; ++t)
#pragma unroll(UNR ? 8 : 1)
                        for (int u = 0; u <= ay + by; ++u) {
                          const double exy = Eab[0][ax][bx][t] * Eab[1][ay][by][u];
                          double sv = 0.0;
#pragma unroll(UNR ? 8 : 1)
                          for (int v = 0; v <= az + bz; ++v) sv = fma(Eab[2][az][bz][v], G[hidx(LAB, t, u, v)], sv);
                          s = fma(exy, sv, s);
                        }
                      acc[ia + NA * (ib + NB * (ic + NC * (id - DLO)))] += s;
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
  }
}

// non-axial normalisation factor of Cartesian component (x,y,z) of a shell with L = x+y+z
// (ShellModule.jl:50-58): sqrt((2L-1)!! / ((2x-1)!!(2y-1)!!(2z-1)!!)).  L<=2: only d_xy,d_xz,d_yz differ (sqrt 3).
__host__ __device__ constexpr double dfact_c(int n) { return n <= 1 ? 1.0 : n * dfact_c(n - 2); }
__device__ __forceinline__ double nonaxial(int L, int x, int y, int z) {
  if (L < 2) return 1.0;
  if (L == 2) return (x == 2 || y == 2 || z == 2) ? 1.0 : 1.7320508075688772;
  return sqrt(dfact_c(2 * L - 1) / (dfact_c(2 * x - 1) * dfact_c(2 * y - 1) * dfact_c(2 * z - 1)));
}

}  // namespace qlb
