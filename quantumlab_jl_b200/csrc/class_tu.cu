// class_tu.cu -- one translation unit per canonical (ab|cd) class: compiled with
//   -DQC_LA= -DQC_LB= -DQC_LC= -DQC_LD=   the four shell angular momenta
//   -DQC_UNR=0|1|2                        flavour of the J/K (mode 0) and tensor (mode 1) kernel:
//                                           0 rolled loops over local-memory arrays
//                                           1 fully unrolled, register resident, one thread per quartet (x ket-d component)
//                                           2 cooperative (eri_coop.cuh): ncart(LC) threads per quartet, R_tuv in shared memory
//                                           3 light (total L <= 2): as 1 with the bra primitives and E^ab staged in shared memory
//   -DQC_DSTEP=n                          ket-d components accumulated per pass (ncart(LD) = single pass)
//   -DQC_MINB=n                           minimum resident CTAs per SM (register cap)
//   -DQC_CPT=n                            cooperative flavour: ket-c components per thread (warps per CTA = ncart(LC)/n)
//   -DQC_SPLIT=0|1                        1: the class is ncart(LD) launches, one per ket-d component, each compiled in
//                                         its own TU (this file again with -DQC_DSEL=k)
// so that the classes build in parallel and each one's policy is a Makefile entry (see DESIGN.md "K4").
//   -DQC_RIONLY=1                         classes that only occur on the RI path (an auxiliary f shell in the bra): mode 2 only
// mode 2 (resolution-of-the-identity scatter) always runs the rolled flavour.
// mode -1 is a query: returns (bra pairs per CTA) | (kernel launches per class launch) << 8 without launching.
#include "qlb_internal.h"
#if QC_UNR == 2
#include "eri_coop.cuh"
#endif

#ifndef QC_MINB
#define QC_MINB 1
#endif

// dynamic shared memory of the register-flavour J/K kernels: TILE_BRA x 2 rows of the density-bound table (int16)
#define QC_DLSMEM(prm) ((prm).dl_smem ? (size_t)TILE_BRA * 2 * (prm).nsh * sizeof(short) : (size_t)0)

#define QC_CAT2(a, b, c, d) launch_qc_##a##b##c##d
#define QC_CAT(a, b, c, d) QC_CAT2(a, b, c, d)
#define QC_DCAT2(a, b, c, d, k) launch_qc_##a##b##c##d##_d##k
#define QC_DCAT(a, b, c, d, k) QC_DCAT2(a, b, c, d, k)

namespace qlb {

#ifndef QC_CPT
#define QC_CPT 1
#endif
#if QC_UNR == 2
// one kernel serves the J/K build (mode 0) and the tensor scatter (mode 1: prm.tensor set by the caller)
template <int DSEL>
static int launch_coop(unsigned grid, cudaStream_t s, const ClassParams &prm) {
  using S = cop::CoopShape<QC_LA, QC_LB, QC_LC, QC_LD, QC_CPT>;
  auto kern = cop::eri_coop_kernel<QC_LA, QC_LB, QC_LC, QC_LD, DSEL, QC_CPT, QC_MINB>;
  static unsigned attr_set = 0;  // one bit per device: the attribute belongs to the device's copy of the kernel
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_set >> dev) & 1u)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set |= 1u << dev;
  }
  kern<<<grid, dim3(32, S::NW), S::SMEM_BYTES, s>>>(prm);
  return (int)cudaGetLastError();
}
#endif

#ifdef QC_DSEL
// ---- one ket-d component of a split class
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, QC_DSEL)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
#if QC_UNR == 2
  return launch_coop<QC_DSEL>(grid, s, prm);
#else
  const dim3 block(TILE_KET, TILE_BRA);
#if QC_UNR
  if (mode == 0) unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, 1, QC_DSEL, QC_MINB><<<grid, block, QC_DLSMEM(prm), s>>>(prm);
  else unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, 1, QC_DSEL, QC_MINB><<<grid, block, 0, s>>>(prm);
#else
  // rolled split classes also split over the ket-c components (gridDim.y): 6x more, 6x shorter threads
  if (mode == 0) rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, 1, QC_DSEL, QC_MINB, true><<<dim3(grid, ncart(QC_LC)), block, QC_DLSMEM(prm), s>>>(prm);
  else rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, 1, QC_DSEL, QC_MINB, false><<<grid, block, 0, s>>>(prm);
#endif
  return (int)cudaGetLastError();
#endif
}
#else
#if QC_SPLIT
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 0)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 1)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 2)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
#if QC_LD >= 2
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 3)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 4)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 5)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
#endif
#endif

int QC_CAT(QC_LA, QC_LB, QC_LC, QC_LD)(int mode, unsigned grid, cudaStream_t s, const ClassParams &prm) {
  const dim3 block(TILE_KET, TILE_BRA);
  if (mode == -1) {
    const int tile_bra = (QC_UNR == 2) ? 1 : TILE_BRA;
    const int nlaunch = QC_SPLIT ? ncart(QC_LD) : 1;
    return tile_bra | (nlaunch << 8);
  }
  if (mode == 0 || mode == 1) {
#if defined(QC_RIONLY) && QC_RIONLY
    return (int)cudaErrorInvalidValue;  // no J/K or tensor kernel exists for this class
#elif QC_SPLIT
    int rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 0)(mode, grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 1)(mode, grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 2)(mode, grid, s, prm);
#if QC_LD >= 2
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 3)(mode, grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 4)(mode, grid, s, prm);
    if (!rc) rc = QC_DCAT(QC_LA, QC_LB, QC_LC, QC_LD, 5)(mode, grid, s, prm);
#endif
    return rc;
#elif QC_UNR == 2
    return launch_coop<0>(grid, s, prm);
#elif QC_UNR == 3
    // light flavour (bra primitives staged in shared memory, ket primitive loop outside)
    if (mode == 0) unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, ncart(QC_LD), -1, QC_MINB, false, true><<<grid, block, QC_DLSMEM(prm), s>>>(prm);
    else unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, ncart(QC_LD), -1, QC_MINB, false, true><<<grid, block, 0, s>>>(prm);
#elif QC_UNR
    if (mode == 0) unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, QC_DSTEP, -1, QC_MINB><<<grid, block, QC_DLSMEM(prm), s>>>(prm);
    else unr::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, QC_DSTEP, -1, QC_MINB><<<grid, block, 0, s>>>(prm);
#else
    if (mode == 0) rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 0, QC_DSTEP, -1, QC_MINB><<<grid, block, QC_DLSMEM(prm), s>>>(prm);
    else rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 1, QC_DSTEP, -1, QC_MINB><<<grid, block, 0, s>>>(prm);
#endif
  } else {
    // RI scatter (mode 2): rolled flavour, TILE_BRA bra pairs per CTA
    rol::eri_class_kernel<QC_LA, QC_LB, QC_LC, QC_LD, 2, ncart(QC_LD)><<<grid, block, 0, s>>>(prm);
  }
  return (int)cudaGetLastError();
}
#endif

}  // namespace qlb
