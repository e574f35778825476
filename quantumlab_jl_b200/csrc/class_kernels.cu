// class_kernels.cu -- instantiation + launch table of the 21 canonical (ab|cd) class kernels (s,p,d).
#include "qlb_internal.h"

namespace qlb {

// UNR / DSTEP policy per class lives here (see DESIGN.md "Kernel K4").
template <int LA, int LB, int LC, int LD>
struct Policy {
  static constexpr int L = LA + LB + LC + LD;
  static constexpr bool unroll = (L <= 2);                 // fully unrolled, register resident
  static constexpr int dstep = ncart(LD);                   // single pass over the ket-d components
};

template <int LA, int LB, int LC, int LD>
static int launch4(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
  using Pol = Policy<LA, LB, LC, LD>;
  const dim3 block(TILE_KET, TILE_BRA);
  if (mode == 0)
    eri_class_kernel<LA, LB, LC, LD, Pol::unroll, 0, Pol::dstep><<<grid, block, 0, s>>>(prm);
  else
    eri_class_kernel<LA, LB, LC, LD, false, 1, ncart(LD)><<<grid, block, 0, s>>>(prm);
  return (int)cudaGetLastError();
}

#define QLB_CASE(X, Y, LA, LB, LC, LD) \
  case (X) * ((X) + 1) / 2 + (Y):      \
    return launch4<LA, LB, LC, LD>(mode, grid, s, prm);

int launch_class_kernel(int X, int Y, int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
  switch (qclass_of(X, Y)) {
    QLB_CASE(0, 0, 0, 0, 0, 0)
    QLB_CASE(1, 0, 1, 0, 0, 0)
    QLB_CASE(1, 1, 1, 0, 1, 0)
    QLB_CASE(2, 0, 1, 1, 0, 0)
    QLB_CASE(2, 1, 1, 1, 1, 0)
    QLB_CASE(2, 2, 1, 1, 1, 1)
    QLB_CASE(3, 0, 2, 0, 0, 0)
    QLB_CASE(3, 1, 2, 0, 1, 0)
    QLB_CASE(3, 2, 2, 0, 1, 1)
    QLB_CASE(3, 3, 2, 0, 2, 0)
    QLB_CASE(4, 0, 2, 1, 0, 0)
    QLB_CASE(4, 1, 2, 1, 1, 0)
    QLB_CASE(4, 2, 2, 1, 1, 1)
    QLB_CASE(4, 3, 2, 1, 2, 0)
    QLB_CASE(4, 4, 2, 1, 2, 1)
    QLB_CASE(5, 0, 2, 2, 0, 0)
    QLB_CASE(5, 1, 2, 2, 1, 0)
    QLB_CASE(5, 2, 2, 2, 1, 1)
    QLB_CASE(5, 3, 2, 2, 2, 0)
    QLB_CASE(5, 4, 2, 2, 2, 1)
    QLB_CASE(5, 5, 2, 2, 2, 2)
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
