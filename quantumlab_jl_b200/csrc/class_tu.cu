// class_tu.cu -- one translation unit per canonical (ab|cd) class: compiled 21 times with
//   -DQC_LA= -DQC_LB= -DQC_LC= -DQC_LD=   the four shell angular momenta
//   -DQC_UNR=0|1                          rolled (local-memory) or fully unrolled (register) flavour of the J/K kernel
//   -DQC_DSTEP=n                          ket-d components accumulated per pass (ncart(LD) = single pass)
// so that the classes build in parallel and each one's policy is a Makefile entry (see DESIGN.md "K4").
#include "qlb_internal.h"

#define QC_CAT2(a, b, c, d) launch_qc_##a##b##c##d
#define QC_CAT(a, b, c, d) QC_CAT2(a, b, c, d)

namespace qlb {

int QC_CAT(QC_LA, QC_LB, QC_LC, QC_LD)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
  const dim3 block(TILE_KET, TILE_BRA);
  if (mode == 0) {
#if QC_UNR
    unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, QC_DSTEP><<<grid, block, 0, s>>>(prm);
#else
    rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, QC_DSTEP><<<grid, block, 0, s>>>(prm);
#endif
  } else {
    rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, ncart(QC_LD)><<<grid, block, 0, s>>>(prm);
  }
  return (int)cudaGetLastError();
}

}  // namespace qlb
