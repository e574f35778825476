// one_electron.cu -- overlap, kinetic and nuclear-attraction matrices on the device.
//
// Replaces computeMatrixOverlap/Kinetic/NuclearAttraction (SpecialMatricesModule.jl:102-177) over
// computeIntegralOverlap (IntegralsModule.jl:138-163), computeIntegralKinetic (:228-273) and
// computeIntegralNuclearAttraction (:306-351).  Same values, evaluated with Hermite expansions instead of the
// reference's closed forms: one thread per ordered shell pair, rolled loops (cold path: runs once per geometry).
#include "qlb_internal.h"

namespace qlb {

constexpr int OE_LMAX = QLB_MAX_L;           // shell L
constexpr int OE_LB = OE_LMAX + 2;           // ket index raised by two for the kinetic operator
constexpr int OE_LT = OE_LMAX + OE_LB;       // max t

__device__ static void boys_rt(int L, double T, const double *__restrict__ tab, double *F) {
  const double e = exp(-T);
  if (T < BOYS_TMAX) {
    const int i = (int)(T * (1.0 / BOYS_DT) + 0.5);
    const double d = (double)i * BOYS_DT - T;
    const double *row = tab + (size_t)L * BOYS_TABLE_DOUBLES + i * BOYS_ROWLEN;
    double f = row[7];
    for (int k = 6; k >= 0; --k) f = fma(f, d, row[k]);  // rows hold F_{L+k}/k!
    F[L] = f;
    for (int m = L; m >= 1; --m) F[m - 1] = (2.0 * T * F[m] + e) / (2 * m - 1);
  } else {
    F[0] = 0.5 * sqrt(QLB_PI / T);
    for (int m = 0; m < L; ++m) F[m + 1] = ((2 * m + 1) * F[m] - e) / (2.0 * T);
  }
}

__device__ static void hermite_E_rt(int la, int lb, double PA, double PB, double oo2p,
                                    double (&E)[OE_LMAX + 1][OE_LB + 1][OE_LT + 1]) {
  for (int i = 0; i <= la; ++i)
    for (int j = 0; j <= lb; ++j)
      for (int t = 0; t <= OE_LT; ++t) E[i][j][t] = 0.0;
  E[0][0][0] = 1.0;
  for (int i = 0; i <= la; ++i)
    for (int j = 0; j <= lb; ++j) {
      if (i == 0 && j == 0) continue;
      for (int t = 0; t <= i + j; ++t) {
        double v;
        if (j == 0) {
          v = PA * E[i - 1][0][t];
          if (t > 0) v += oo2p * E[i - 1][0][t - 1];
          if (t + 1 <= i - 1) v += (t + 1) * E[i - 1][0][t + 1];
        } else {
          v = PB * E[i][j - 1][t];
          if (t > 0) v += oo2p * E[i][j - 1][t - 1];
          if (t + 1 <= i + j - 1) v += (t + 1) * E[i][j - 1][t + 1];
        }
        E[i][j][t] = v;
      }
    }
}

struct OneEParams {
  const double *xyz; const int *L; const int *poff; const int *boff; const double *exps; const double *coefs;
  const double *boys;
  int nsh, nbf, which, natoms;
  const double *atoms;  // natoms x (x,y,z,Z)
  double *out;
};

__global__ void __launch_bounds__(64) one_electron_kernel(const __grid_constant__ OneEParams prm) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= prm.nsh * prm.nsh) return;
  const int si = idx % prm.nsh, sj = idx / prm.nsh;
  const int la = prm.L[si], lb = prm.L[sj];
  const int na = ncart(la), nb = ncart(lb);
  const double Ax = prm.xyz[3 * si], Ay = prm.xyz[3 * si + 1], Az = prm.xyz[3 * si + 2];
  const double Bx = prm.xyz[3 * sj], By = prm.xyz[3 * sj + 1], Bz = prm.xyz[3 * sj + 2];
  const double ab2 = (Ax - Bx) * (Ax - Bx) + (Ay - By) * (Ay - By) + (Az - Bz) * (Az - Bz);
  double blk[ncart(OE_LMAX) * ncart(OE_LMAX)];
  for (int k = 0; k < na * nb; ++k) blk[k] = 0.0;
  double E[3][OE_LMAX + 1][OE_LB + 1][OE_LT + 1];
  const int lbx = lb + (prm.which == 1 ? 2 : 0);
  for (int ia = prm.poff[si]; ia < prm.poff[si + 1]; ++ia)
    for (int ib = prm.poff[sj]; ib < prm.poff[sj + 1]; ++ib) {
      const double al = prm.exps[ia], be = prm.exps[ib], p = al + be, ip = 1.0 / p;
      const double Px = (al * Ax + be * Bx) * ip, Py = (al * Ay + be * By) * ip, Pz = (al * Az + be * Bz) * ip;
      const double cc = prm.coefs[ia] * prm.coefs[ib] * exp(-al * be * ip * ab2);
      hermite_E_rt(la, lbx, Px - Ax, Px - Bx, 0.5 * ip, E[0]);
      hermite_E_rt(la, lbx, Py - Ay, Py - By, 0.5 * ip, E[1]);
      hermite_E_rt(la, lbx, Pz - Az, Pz - Bz, 0.5 * ip, E[2]);
      if (prm.which < 2) {
        const double s0 = sqrt(QLB_PI * ip);  // 1-D overlap of the s product
        int kb = 0;
        for (int bx = lb; bx >= 0; --bx)
          for (int by = lb - bx; by >= 0; --by, ++kb) {
            const int bz = lb - bx - by;
            int ka = 0;
            for (int ax = la; ax >= 0; --ax)
              for (int ay = la - ax; ay >= 0; --ay, ++ka) {
                const int az = la - ax - ay;
                const double Sx = E[0][ax][bx][0] * s0, Sy = E[1][ay][by][0] * s0, Sz = E[2][az][bz][0] * s0;
                double v;
                if (prm.which == 0) v = Sx * Sy * Sz;
                else {
                  // T = -1/2 <a| d2/dx2 + .. |b>;  per axis: D = 4 be^2 S(j+2) - 2 be (2j+1) S(j) + j(j-1) S(j-2)
                  const int a3[3] = {ax, ay, az}, b3[3] = {bx, by, bz};
                  double S1[3] = {Sx, Sy, Sz}, D[3];
                  for (int d = 0; d < 3; ++d) {
                    const int i = a3[d], j = b3[d];
                    double t = 4.0 * be * be * E[d][i][j + 2][0] - 2.0 * be * (2 * j + 1) * E[d][i][j][0];
                    if (j >= 2) t += (double)(j * (j - 1)) * E[d][i][j - 2][0];
                    D[d] = t * s0;
                  }
                  v = -0.5 * (D[0] * S1[1] * S1[2] + S1[0] * D[1] * S1[2] + S1[0] * S1[1] * D[2]);
                }
                blk[ka + na * kb] += cc * v;
              }
          }
      } else {
        const int Lt = la + lb;
        double R[nherm(2 * OE_LMAX)], F[2 * OE_LMAX + 1];
        for (int at = 0; at < prm.natoms; ++at) {
          const double X = Px - prm.atoms[4 * at], Y = Py - prm.atoms[4 * at + 1], Zc = Pz - prm.atoms[4 * at + 2];
          const double pref = -prm.atoms[4 * at + 3] * 2.0 * QLB_PI * ip * cc;
          boys_rt(Lt, p * (X * X + Y * Y + Zc * Zc), prm.boys, F);
          // in-place Hermite Coulomb recursion (same scheme as hermite_R), runtime order Lt, layout of order 2*OE_LMAX
          constexpr int LL = 2 * OE_LMAX;
          double s = 1.0, Fs[LL + 1];
          for (int n = 0; n <= Lt; ++n) { Fs[n] = F[n] * s * pref; s *= -2.0 * p; }
          R[0] = Fs[Lt];
          for (int n = Lt - 1; n >= 0; --n) {
            for (int k = Lt - n; k >= 1; --k)
              for (int t = k; t >= 0; --t)
                for (int u = k - t; u >= 0; --u) {
                  const int v = k - t - u;
                  double val;
                  if (t > 0) { val = X * R[hidx(LL, t - 1, u, v)]; if (t > 1) val += (t - 1) * R[hidx(LL, t - 2, u, v)]; }
                  else if (u > 0) { val = Y * R[hidx(LL, t, u - 1, v)]; if (u > 1) val += (u - 1) * R[hidx(LL, t, u - 2, v)]; }
                  else { val = Zc * R[hidx(LL, t, u, v - 1)]; if (v > 1) val += (v - 1) * R[hidx(LL, t, u, v - 2)]; }
                  R[hidx(LL, t, u, v)] = val;
                }
            R[0] = Fs[n];
          }
          int kb = 0;
          for (int bx = lb; bx >= 0; --bx)
            for (int by = lb - bx; by >= 0; --by, ++kb) {
              const int bz = lb - bx - by;
              int ka = 0;
              for (int ax = la; ax >= 0; --ax)
                for (int ay = la - ax; ay >= 0; --ay, ++ka) {
                  const int az = la - ax - ay;
                  double v = 0.0;
                  for (int t = 0; t <= ax + bx; ++t)
                    for (int u = 0; u <= ay + by; ++u)
                      for (int w = 0; w <= az + bz; ++w)
                        v += E[0][ax][bx][t] * E[1][ay][by][u] * E[2][az][bz][w] * R[hidx(LL, t, u, w)];
                  blk[ka + na * kb] += v;
                }
            }
        }
      }
    }
  // non-axial factors, scatter
  int kb = 0;
  for (int bx = lb; bx >= 0; --bx)
    for (int by = lb - bx; by >= 0; --by, ++kb) {
      const double fb = nonaxial(lb, bx, by, lb - bx - by);
      int ka = 0;
      for (int ax = la; ax >= 0; --ax)
        for (int ay = la - ax; ay >= 0; --ay, ++ka)
          prm.out[(prm.boff[si] + ka) + (size_t)prm.nbf * (prm.boff[sj] + kb)] = blk[ka + na * kb] * fb * nonaxial(la, ax, ay, la - ax - ay);
    }
}

int one_electron_device(qlb_basis *b, int which, int natoms, const double *d_atoms, double *d_out, cudaStream_t s) {
  OneEParams prm;
  prm.xyz = b->d_xyz; prm.L = b->d_L; prm.poff = b->d_poff; prm.boff = b->d_boff; prm.exps = b->d_exps; prm.coefs = b->d_coefs;
  prm.boys = engine().d_boys;
  prm.nsh = b->nsh; prm.nbf = b->nbf; prm.which = which; prm.natoms = natoms; prm.atoms = d_atoms; prm.out = d_out;
  const int n = b->nsh * b->nsh;
  one_electron_kernel<<<(n + 63) / 64, 64, 0, s>>>(prm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "one_electron_kernel");
  return QLB_OK;
}

}  // namespace qlb
