// class_kernels.cu -- launch table of the 21 canonical (ab|cd) class kernels (s,p,d) + Schwarz kernels + flop model.
#include "qlb_internal.h"

namespace qlb {

// one launcher per class, each defined in its own translation unit (class_tu.cu compiled with -DQC_*)
int launch_qc_0000(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_1000(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_1010(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_1100(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_1110(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_1111(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2000(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2010(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2011(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2020(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2100(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2110(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2111(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2120(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2121(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2200(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2210(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2211(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2220(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2221(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_2222(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
// (auxiliary f shell, unit s) pairs of the RI path: bra class 6 = (fs); rolled MODE-2 kernels only
int launch_qc_3000(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_3010(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_3011(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_3020(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_3021(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_3022(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_qc_3030(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);

int launch_class_kernel(int X, int Y, int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
  switch (qclass_of(X, Y)) {
    case 0: return launch_qc_0000(mode, grid, s, prm);  // (ss|ss)
    case 1: return launch_qc_1000(mode, grid, s, prm);  // (ps|ss)
    case 2: return launch_qc_1010(mode, grid, s, prm);  // (ps|ps)
    case 3: return launch_qc_1100(mode, grid, s, prm);  // (pp|ss)
    case 4: return launch_qc_1110(mode, grid, s, prm);  // (pp|ps)
    case 5: return launch_qc_1111(mode, grid, s, prm);  // (pp|pp)
    case 6: return launch_qc_2000(mode, grid, s, prm);  // (ds|ss)
    case 7: return launch_qc_2010(mode, grid, s, prm);  // (ds|ps)
    case 8: return launch_qc_2011(mode, grid, s, prm);  // (ds|pp)
    case 9: return launch_qc_2020(mode, grid, s, prm);  // (ds|ds)
    case 10: return launch_qc_2100(mode, grid, s, prm);  // (dp|ss)
    case 11: return launch_qc_2110(mode, grid, s, prm);  // (dp|ps)
    case 12: return launch_qc_2111(mode, grid, s, prm);  // (dp|pp)
    case 13: return launch_qc_2120(mode, grid, s, prm);  // (dp|ds)
    case 14: return launch_qc_2121(mode, grid, s, prm);  // (dp|dp)
    case 15: return launch_qc_2200(mode, grid, s, prm);  // (dd|ss)
    case 16: return launch_qc_2210(mode, grid, s, prm);  // (dd|ps)
    case 17: return launch_qc_2211(mode, grid, s, prm);  // (dd|pp)
    case 18: return launch_qc_2220(mode, grid, s, prm);  // (dd|ds)
    case 19: return launch_qc_2221(mode, grid, s, prm);  // (dd|dp)
    case 20: return launch_qc_2222(mode, grid, s, prm);  // (dd|dd)
    case 21: return launch_qc_3000(mode, grid, s, prm);  // (fs|ss)  -- RI only from here
    case 22: return launch_qc_3010(mode, grid, s, prm);  // (fs|ps)
    case 23: return launch_qc_3011(mode, grid, s, prm);  // (fs|pp)
    case 24: return launch_qc_3020(mode, grid, s, prm);  // (fs|ds)
    case 25: return launch_qc_3021(mode, grid, s, prm);  // (fs|dp)
    case 26: return launch_qc_3022(mode, grid, s, prm);  // (fs|dd)
    case 27: return launch_qc_3030(mode, grid, s, prm);  // (fs|fs)
  }
  return (int)cudaErrorInvalidValue;
}

int launch_schwarz_kernel(int X, unsigned grid, cudaStream_t s, const SchwarzParams &prm) {
  switch (X) {
    case 0: schwarz_kernel<0, 0><<<grid, 64, 0, s>>>(prm); break;
    case 1: schwarz_kernel<1, 0><<<grid, 64, 0, s>>>(prm); break;
    case 2: schwarz_kernel<1, 1><<<grid, 64, 0, s>>>(prm); break;
    case 3: schwarz_kernel<2, 0><<<grid, 64, 0, s>>>(prm); break;
    case 4: schwarz_kernel<2, 1><<<grid, 64, 0, s>>>(prm); break;
    case 5: schwarz_kernel<2, 2><<<grid, 64, 0, s>>>(prm); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

// minimum McMurchie-Davidson flop count per primitive quartet and digestion flops per contracted quartet
static double binom4(int n) { return n < 4 ? 0.0 : (double)n * (n - 1) * (n - 2) * (n - 3) / 24.0; }
void class_flops(int X, int Y, double &f_prim, double &f_digest) {
  int la, lb, lc, ld;
  class_L(X, la, lb);
  class_L(Y, lc, ld);
  const int Lab = la + lb, Lcd = lc + ld, L = Lab + Lcd;
  const double nCab = ncart(la) * ncart(lb), nCcd = ncart(lc) * ncart(ld);
  f_prim = (30.0 + 3.0 * L) + 2.0 * binom4(L + 4) + 2.0 * nherm(Lab) * nCcd * (nherm(Lcd) + nCab);
  f_digest = 12.0 * nCab * nCcd;
}

}  // namespace qlb
