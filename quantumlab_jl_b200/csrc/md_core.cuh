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
// One table per total angular momentum L (a class kernel needs exactly one): rows T_i = i*BOYS_DT, each row
// {F_{L+k}(T_i)/k!, k = 0..7, exp(-T_i), pad} = 80 B, read as 16-byte vectors.  29 KB per L, L1 resident (a shared-memory
// copy was measured slower: the carve-out shrinks L1 for the scattered pair-record / density loads).
// F_L(T) by an 8-term Taylor expansion about the nearest row, lower orders by
// downward recursion with exp(-T) = exp(-T_i) * poly(T_i - T); beyond the table the asymptotic F_0 and upward
// recursion (without the exp(-T) term, which is below rounding there).  Accurate to ~2e-15 relative.
constexpr int BOYS_ROWLEN = 10;
constexpr int BOYS_LMAX = 8;            // (dd|dd)
constexpr double BOYS_DT = 0.1;
constexpr double BOYS_TMAX = 36.0;
constexpr int BOYS_NROW = 362;          // rows 0..361 (T up to 36.1)
constexpr int BOYS_TABLE_DOUBLES = BOYS_NROW * BOYS_ROWLEN;

// tab: the table of THIS L.  Rows hold the Taylor coefficients F_{L+k}(T_i)/k! (k = 0..7) so that the expansion is a bare
// Horner chain (ncu round 2: the seven d/k multiplications were a third of the FP64 instructions of the table path).
template <int L>
__device__ __forceinline__ void boys_table(double T, const double *tab, double *F) {
  const int i = (int)(T * (1.0 / BOYS_DT) + 0.5);
  const double d = (double)i * BOYS_DT - T;  // T_i - T, |d| <= DT/2
  const double2 *row = reinterpret_cast<const double2 *>(tab + i * BOYS_ROWLEN);
  const double2 r0 = __ldg(row), r1 = __ldg(row + 1), r2 = __ldg(row + 2), r3 = __ldg(row + 3);
  double f = r3.y;
  f = fma(f, d, r3.x); f = fma(f, d, r2.y); f = fma(f, d, r2.x);
  f = fma(f, d, r1.y); f = fma(f, d, r1.x); f = fma(f, d, r0.y);
  f = fma(f, d, r0.x);
  F[L] = f;
  if (L > 0) {
    // exp(-T) = exp(-T_i) * exp(d): tabulated exp(-T_i) times a degree-8 Taylor polynomial (|d| <= 0.05)
    double e = 1.0 / 40320.0;
    e = fma(e, d, 1.0 / 5040.0); e = fma(e, d, 1.0 / 720.0); e = fma(e, d, 1.0 / 120.0); e = fma(e, d, 1.0 / 24.0);
    e = fma(e, d, 1.0 / 6.0); e = fma(e, d, 0.5); e = fma(e, d, 1.0); e = fma(e, d, 1.0);
    e *= __ldg(row + 4).x;
    const double t2 = 2.0 * T;
#pragma unroll
    for (int m = L; m >= 1; --m) F[m - 1] = fma(t2, F[m], e) * (1.0 / (2 * m - 1));
  }
}

template <int L>
__device__ __forceinline__ void boys_array(double T, const double *tab, double *F) {
  if (T < BOYS_TMAX) {
    boys_table<L>(T, tab, F);
  } else if constexpr (L <= 2) {
    // T >= 36, low orders: F_0 = sqrt(pi/T)/2 and the upward recursion WITHOUT its exp(-T) term: exp(-36) = 2.3e-16 is
    // below the rounding error of the terms it is added to (relative 1.6e-15 for F_1, 4e-14 for F_2 at T = 36, smaller
    // beyond; for higher orders the term matters -- F_8(36) ~ 4e-10 -- so they keep it).  ncu: most primitive quartets of
    // a screened build take this branch, and with exp() it cost more instructions than the table branch.
    const double rs = rsqrt(T);
    F[0] = 0.88622692545275801365 * rs;  // sqrt(pi)/2 / sqrt(T)
    if (L > 0) {
      const double i2t = 0.5 * (rs * rs);
#pragma unroll
      for (int m = 0; m < L; ++m) F[m + 1] = (double)(2 * m + 1) * F[m] * i2t;
    }
  } else {
    const double it = 1.0 / T;
    F[0] = 0.5 * sqrt(QLB_PI * it);
    const double e = exp(-T);
    const double i2t = 0.5 * it;
#pragma unroll
    for (int m = 0; m < L; ++m) F[m + 1] = fma((double)(2 * m + 1), F[m], -e) * i2t;
  }
}

// ------------------------------------------------------------------ primitive-pair records (SoA in HBM)
// one 48-byte record per primitive pair: {p, Px}, {Py, Pz}, {c_a c_b exp(-a b/p |AB|^2) / p, 1/(2p)}
struct PrimPairs {
  const double2 *__restrict__ rec;
};

struct PairRef {  // what one thread needs to know about one shell pair
  double Ax, Ay, Az, Bx, By, Bz;
  int off, n;     // double2 index of the first primitive record (tile layout, see eri_body.inc); primitive pairs
};

// non-axial normalisation factor of Cartesian component (x,y,z) of a shell with L = x+y+z
// (ShellModule.jl:50-58): sqrt((2L-1)!! / ((2x-1)!!(2y-1)!!(2z-1)!!)).  L<=2: only d_xy,d_xz,d_yz differ (sqrt 3).
__host__ __device__ constexpr double dfact_c(int n) { return n <= 1 ? 1.0 : n * dfact_c(n - 2); }
__device__ __forceinline__ double nonaxial(int L, int x, int y, int z) {
  if (L < 2) return 1.0;
  if (L == 2) return (x == 2 || y == 2 || z == 2) ? 1.0 : 1.7320508075688772;
  return sqrt(dfact_c(2 * L - 1) / (dfact_c(2 * x - 1) * dfact_c(2 * y - 1) * dfact_c(2 * z - 1)));
}

}  // namespace qlb
