// class_tu.cu -- one translation unit per canonical (ab|cd) class: compiled with
//   -DQC_LA= -DQC_LB= -DQC_LC= -DQC_LD=   the four shell angular momenta
//   -DQC_UNR=0|1                          rolled (local-memory) or fully unrolled (register) flavour of the J/K kernel
//   -DQC_DSTEP=n                          ket-d components accumulated per pass (ncart(LD) = single pass)
//   -DQC_MINB=n                           minimum resident CTAs per SM (register cap = 65536/(128 n))
//   -DQC_SPLIT=0|1                        1: the J/K build of this class is ncart(LD) launches, one per ket-d component,
//                                         each compiled in its own TU (this file again with -DQC_DSEL=k)
// so that the classes build in parallel and each one's policy is a Makefile entry (see DESIGN.md "K4").
#include "qlb_internal.h"

#ifndef QC_MINB
#define QC_MINB 1
#endif

#define QC_CAT2(a, b, c, d) launch_qc_##a##b##c##d
#define QC_CAT(a, b, c, d) QC_CAT2(a, b, c, d)
#define QC_DCAT2(a, b, c, d, k) launch_qc_##a##b##c##d##_d##k
#define QC_DCAT(a, b, c, d, k) QC_DCAT2(a, b, c, d, k)

namespace qlb {

#ifdef QC_DSEL
// ---- one ket-d component of a split class
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, QC_DSEL)(unsigned grid, cudaStream_t s, const ClassParams &prm) {
  const dim3 block(TILE_KET, TILE_BRA);
#if QC_UNR
  unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, 1, QC_DSEL, QC_MINB><<<grid, block, 0, s>>>(prm);
#else
  // rolled split classes also split over the ket-c components (gridDim.y): 6x more, 6x shorter threads
  rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, 1, QC_DSEL, QC_MINB, true><<<dim3(grid, ncart(QC_LC)), block, 0, s>>>(prm);
#endif
  return (int)cudaGetLastError();
}
#else
#if QC_SPLIT
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 0)(unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 1)(unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 2)(unsigned grid, cudaStream_t s, const ClassParams &prm);
#if QC_LD >= 2
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 3)(unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 4)(unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 5)(unsigned grid, cudaStream_t s, const ClassParams &prm);
#endif
#endif

int QC_CAT(QC_LA, QC_LB, QC_LC, QC_LD)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
  const dim3 block(TILE_KET, TILE_BRA);
  if (mode == 0) {
#if QC_SPLIT
    int rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 0)(grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 1)(grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 2)(grid, s, prm);
#if QC_LD >= 2
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 3)(grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 4)(grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 5)(grid, s, prm);
#endif
    return rc;
#elif QC_UNR
    unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, QC_DSTEP, -1, QC_MINB><<<grid, block, 0, s>>>(prm);
#else
    rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, QC_DSTEP, -1, QC_MINB><<<grid, block, 0, s>>>(prm);
#endif
  } else if (mode == 1) {
    rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, ncart(QC_LD)><<<grid, block, 0, s>>>(prm);
  } else {
    rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 2, ncart(QC_LD)><<<grid, block, 0, s>>>(prm);
  }
  return (int)cudaGetLastError();
}
#endif

}  // namespace qlb
