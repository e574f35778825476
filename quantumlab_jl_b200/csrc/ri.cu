// ri.cu -- resolution of the identity in the Coulomb metric.
//
// Replaces ResolutionIdentityModule.jl: computeMatrixResolutionOfTheIdentityCoulombMetric (:41-50),
// computeTensorElectronRepulsionIntegralsRICoulomb (:54-58) and computeMatrixExchangeRIK (:69-78).  The reference
// forms (mu nu|P) and (P|Q) through its 4-centre ERI with the unit s function `identityCGF`
// (BasisFunctionsModule.jl:50-51), takes B = (mu nu|P) (P|Q)^-1/2, rebuilds the FULL N^4 tensor B B^T and contracts it.
// Here: the same 3- and 2-index integrals come from the class kernels (MODE 2, an (auxiliary shell, unit s) pair on
// one side), the metric power from cuSOLVER (Dsyevd), and every contraction is a dense FP64 GEMM/GEMV in cuBLAS
// (DMMA tensor cores on B200); the N^4 tensor is never formed:
//     J = B (B^T vec P)                      2 GEMV
//     K = sum_Q B^Q P B^Q                    strided-batched GEMM (X^Q = B^Q P) + one GEMM over (sigma,Q) per chunk of Q
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <string>
#include <vector>

#include "qlb_internal.h"

using namespace qlb;

struct qlb_ri {
  qlb_basis *comb = nullptr;  // orbital shells + auxiliary shells + unit s function
  int N = 0, Naux = 0, aux_fn0 = 0;
  double *d_B = nullptr;      // N*N*Naux, column-major [mu, nu, P]
  double *d_P = nullptr, *d_JK = nullptr, *d_v = nullptr, *d_X = nullptr;
  int qchunk = 0;
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
};

namespace {

__global__ void scale_columns_kernel(int n, const double *__restrict__ w, double *V) {
  // V[:, j] *= w[j]^-1/4 ... see metric_inverse_sqrt: V <- V diag(lambda^-1/4) so that V V^T = M^-1/2
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n) V[i + (size_t)n * j] *= w[j];
}

int blas_fail(cublasStatus_t st, const char *what) {
  set_error(std::string("cuBLAS error ") + std::to_string((int)st) + " in " + what);
  return QLB_ERR_CUDA;
}
int solver_fail(cusolverStatus_t st, const char *what) {
  set_error(std::string("cuSOLVER error ") + std::to_string((int)st) + " in " + what);
  return QLB_ERR_CUDA;
}

// launch MODE-2 class kernels: bra group/class vs ket group/class (canonical order handled by the caller)
int ri_launch(qlb_basis *b, ClassParams &prm, int X, const int *xoff, int Y, const int *yoff, int ri_mode, cudaStream_t s) {
  prm.bra_off = xoff[X]; prm.bra_n = xoff[X + 1] - xoff[X];
  prm.ket_off = yoff[Y]; prm.ket_n = yoff[Y + 1] - yoff[Y];
  if (prm.bra_n == 0 || prm.ket_n == 0) return QLB_OK;
  prm.same = 0;
  prm.ri_mode = ri_mode;
  prm.nbx = (prm.ket_n + TILE_KET - 1) / TILE_KET;
  prm.ysplit = prm.nbx;
  const int64_t grid = (int64_t)((prm.bra_n + TILE_BRA - 1) / TILE_BRA) * prm.nbx;
  if (int ce = launch_class_kernel(X, Y, 2, (unsigned)grid, s, prm)) return cuda_fail((cudaError_t)ce, "eri_class_kernel(ri)");
  return QLB_OK;
}

}  // namespace

extern "C" int qlb_ri_destroy(qlb_ri *r) {
  if (!r) return QLB_OK;
  if (engine().ready) cudaSetDevice(engine().device);
  void *ptrs[] = {r->d_B, r->d_P, r->d_JK, r->d_v, r->d_X};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (r->blas) cublasDestroy(r->blas);
  if (r->solver) cusolverDnDestroy(r->solver);
  if (r->comb) qlb_basis_destroy(r->comb);
  delete r;
  return QLB_OK;
}

extern "C" int qlb_ri_naux(const qlb_ri *r, int *naux) {
  if (!r || !naux) return QLB_ERR_ARG;
  *naux = r->Naux;
  return QLB_OK;
}

extern "C" int qlb_ri_create(qlb_basis *orb, int naux_sh, const double *centers, const int32_t *L, const int32_t *nprim,
                             const double *exps, const double *coefs, qlb_ri **out) {
  Engine &e = engine();
  if (!e.ready) {
    set_error("qlb_init has not been called");
    return QLB_ERR_STATE;
  }
  if (!orb || !out || naux_sh <= 0 || !centers || !L || !nprim || !exps || !coefs) {
    set_error("qlb_ri_create: bad argument");
    return QLB_ERR_ARG;
  }
  QLB_CUDA(cudaSetDevice(e.device));
  // combined shell list: orbital | auxiliary | unit s function (exponent 0, coefficient 1, at the origin)
  const int nsh = orb->nsh + naux_sh + 1;
  std::vector<double> xyz(orb->h_xyz), ex(orb->h_exps), co(orb->h_coefs);
  std::vector<int32_t> LL(orb->h_L.begin(), orb->h_L.end()), np(orb->h_nprim.begin(), orb->h_nprim.end());
  int naux_prim = 0;
  for (int i = 0; i < naux_sh; ++i) naux_prim += nprim[i];
  xyz.insert(xyz.end(), centers, centers + 3 * (size_t)naux_sh);
  LL.insert(LL.end(), L, L + naux_sh);
  np.insert(np.end(), nprim, nprim + naux_sh);
  ex.insert(ex.end(), exps, exps + naux_prim);
  co.insert(co.end(), coefs, coefs + naux_prim);
  xyz.insert(xyz.end(), {0.0, 0.0, 0.0});
  LL.push_back(0);
  np.push_back(1);
  ex.push_back(0.0);
  co.push_back(1.0);
  qlb_ri *r = new qlb_ri();
  int rc = basis_create_internal(nsh, xyz.data(), LL.data(), np.data(), ex.data(), co.data(), orb->nsh, &r->comb);
  if (rc) {
    delete r;
    return rc;
  }
  qlb_basis *b = r->comb;
  r->N = orb->nbf;
  r->aux_fn0 = b->h_boff[orb->nsh];
  r->Naux = b->h_boff[nsh - 1] - r->aux_fn0;
  const size_t N = r->N, NX = r->Naux, N2 = N * N;
  cudaError_t ce;
  double *d_B3 = nullptr, *d_M = nullptr, *d_w = nullptr, *d_Mh = nullptr, *d_work = nullptr;
  int *d_info = nullptr;
  auto cleanup = [&]() {
    void *ptrs[] = {d_B3, d_M, d_w, d_Mh, d_work, d_info};
    for (void *p : ptrs)
      if (p) cudaFree(p);
  };
#define RI_CUDA(call)                        \
  do {                                       \
    ce = (call);                             \
    if (ce != cudaSuccess) {                 \
      cleanup();                             \
      qlb_ri_destroy(r);                     \
      return cuda_fail(ce, #call);           \
    }                                        \
  } while (0)
  RI_CUDA(cudaMalloc((void **)&d_B3, N2 * NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_B, N2 * NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_M, NX * NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_Mh, NX * NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_w, NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_info, sizeof(int)));
  RI_CUDA(cudaMalloc((void **)&r->d_P, N2 * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_JK, 2 * N2 * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_v, NX * sizeof(double)));
  // chunk of auxiliary functions for the exchange contraction: X = B^Q P for qchunk Q's at a time (<= 1 GiB scratch)
  r->qchunk = (int)std::max<size_t>(1, std::min<size_t>(NX, ((size_t)1 << 30) / (N2 * sizeof(double))));
  RI_CUDA(cudaMalloc((void **)&r->d_X, N2 * r->qchunk * sizeof(double)));
  RI_CUDA(cudaMemsetAsync(d_B3, 0, N2 * NX * sizeof(double), e.stream));
  RI_CUDA(cudaMemsetAsync(d_M, 0, NX * NX * sizeof(double), e.stream));
  // ---- 3-index (mu nu|P) and 2-index (P|Q) integrals from the class kernels
  ClassParams prm;
  base_class_params(b, prm);
  prm.ri_n = (int)N; prm.ri_naux = (int)NX; prm.ri_aux_fn0 = r->aux_fn0;
  for (int X = 0; X < NCLASS && rc == QLB_OK; ++X)
    for (int Y = 0; Y < NCLASS && rc == QLB_OK; ++Y) {
      // X: orbital pair class, Y: (aux, unit) pair class.  Kernels exist for bra class >= ket class.
      prm.tensor = d_B3;
      if (X >= Y) rc = ri_launch(b, prm, X, b->class_off, Y, b->ri_class_off, 0, e.stream);
      else rc = ri_launch(b, prm, Y, b->ri_class_off, X, b->class_off, 1, e.stream);
    }
  for (int Y1 = 0; Y1 < NCLASS && rc == QLB_OK; ++Y1)
    for (int Y2 = 0; Y2 <= Y1 && rc == QLB_OK; ++Y2) {
      prm.tensor = d_M;
      rc = ri_launch(b, prm, Y1, b->ri_class_off, Y2, b->ri_class_off, 2, e.stream);
    }
  if (rc) {
    cleanup();
    qlb_ri_destroy(r);
    return rc;
  }
  // ---- (P|Q)^-1/2 = V diag(lambda^-1/2) V^T  (the reference: sqrt(inv(C)), ResolutionIdentityModule.jl:49)
  cublasStatus_t bs = cublasCreate(&r->blas);
  cusolverStatus_t ss = cusolverDnCreate(&r->solver);
  if (bs != CUBLAS_STATUS_SUCCESS || ss != CUSOLVER_STATUS_SUCCESS) {
    cleanup();
    qlb_ri_destroy(r);
    return bs != CUBLAS_STATUS_SUCCESS ? blas_fail(bs, "cublasCreate") : solver_fail(ss, "cusolverDnCreate");
  }
  cublasSetStream(r->blas, e.stream);
  cusolverDnSetStream(r->solver, e.stream);
  int lwork = 0;
  ss = cusolverDnDsyevd_bufferSize(r->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)NX, d_M, (int)NX, d_w, &lwork);
  if (ss == CUSOLVER_STATUS_SUCCESS) {
    ce = cudaMalloc((void **)&d_work, std::max(lwork, 1) * sizeof(double));
    if (ce != cudaSuccess) { cleanup(); qlb_ri_destroy(r); return cuda_fail(ce, "cudaMalloc(syevd work)"); }
    ss = cusolverDnDsyevd(r->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)NX, d_M, (int)NX, d_w, d_work, lwork, d_info);
  }
  if (ss != CUSOLVER_STATUS_SUCCESS) {
    cleanup();
    qlb_ri_destroy(r);
    return solver_fail(ss, "cusolverDnDsyevd");
  }
  std::vector<double> w(NX);
  int info = 0;
  RI_CUDA(cudaMemcpyAsync(w.data(), d_w, NX * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  RI_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  RI_CUDA(cudaStreamSynchronize(e.stream));
  if (info != 0 || w[0] <= 0.0) {
    cleanup();
    qlb_ri_destroy(r);
    set_error("qlb_ri_create: the Coulomb metric (P|Q) is not positive definite (linearly dependent auxiliary basis?)");
    return QLB_ERR_ARG;
  }
  for (size_t j = 0; j < NX; ++j) w[j] = pow(w[j], -0.25);  // V <- V diag(lambda^-1/4); V V^T = M^-1/2
  RI_CUDA(cudaMemcpyAsync(d_w, w.data(), NX * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  scale_columns_kernel<<<dim3((unsigned)((NX + 127) / 128), (unsigned)NX), 128, 0, e.stream>>>((int)NX, d_w, d_M);
  const double one = 1.0, zero = 0.0;
  bs = cublasDgemm(r->blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)NX, (int)NX, (int)NX, &one, d_M, (int)NX, d_M, (int)NX, &zero, d_Mh, (int)NX);
  // B = B3 * M^-1/2   (N^2 x Naux) = (N^2 x Naux)(Naux x Naux)
  if (bs == CUBLAS_STATUS_SUCCESS)
    bs = cublasDgemm(r->blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)N2, (int)NX, (int)NX, &one, d_B3, (int)N2, d_Mh, (int)NX, &zero, r->d_B, (int)N2);
  if (bs != CUBLAS_STATUS_SUCCESS) {
    cleanup();
    qlb_ri_destroy(r);
    return blas_fail(bs, "cublasDgemm(B)");
  }
  RI_CUDA(cudaStreamSynchronize(e.stream));
  cleanup();
#undef RI_CUDA
  *out = r;
  return QLB_OK;
}

extern "C" int qlb_ri_b_tensor(qlb_ri *r, double *B) {
  if (!r || !B) return QLB_ERR_ARG;
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t n = (size_t)r->N * r->N * r->Naux;
  QLB_CUDA(cudaMemcpyAsync(B, r->d_B, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  return QLB_OK;
}

// ERI_RI[mu,nu,la,si] = sum_Q B[mu,nu,Q] B[la,si,Q]  (computeTensorElectronRepulsionIntegralsRICoulomb, :54-58; tests only)
extern "C" int qlb_ri_eri_tensor(qlb_ri *r, double *out) {
  if (!r || !out) return QLB_ERR_ARG;
  if (r->N > 64) {
    set_error("qlb_ri_eri_tensor: N > 64; the N^4 tensor is only materialised for test-sized systems");
    return QLB_ERR_ARG;
  }
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)r->N * r->N;
  double *dT = nullptr;
  QLB_CUDA(cudaMalloc((void **)&dT, N2 * N2 * sizeof(double)));
  const double one = 1.0, zero = 0.0;
  cublasStatus_t bs = cublasDgemm(r->blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)N2, (int)N2, r->Naux, &one, r->d_B, (int)N2, r->d_B,
                                  (int)N2, &zero, dT, (int)N2);
  int rc = QLB_OK;
  if (bs != CUBLAS_STATUS_SUCCESS) rc = blas_fail(bs, "cublasDgemm(B B^T)");
  else {
    cudaError_t ce = cudaMemcpyAsync(out, dT, N2 * N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "RI tensor D2H");
  }
  cudaFree(dT);
  return rc;
}

extern "C" int qlb_ri_build_jk(qlb_ri *r, const double *P, double *J, double *K, double *ms) {
  if (!r || !P || (!J && !K)) {
    set_error("qlb_ri_build_jk: null pointer");
    return QLB_ERR_ARG;
  }
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const int N = r->N, NX = r->Naux;
  const size_t N2 = (size_t)N * N;
  const double one = 1.0, zero = 0.0;
  cudaEvent_t ev0 = r->comb->ev[0], ev1 = r->comb->ev[2];
  QLB_CUDA(cudaMemcpyAsync(r->d_P, P, N2 * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  cudaEventRecord(ev0, e.stream);
  cublasStatus_t bs = CUBLAS_STATUS_SUCCESS;
  double *dJ = r->d_JK, *dK = r->d_JK + N2;
  if (J) {
    // v_Q = sum_{la si} B[la si, Q] P[la si];  J[mu nu] = sum_Q B[mu nu, Q] v_Q
    bs = cublasDgemv(r->blas, CUBLAS_OP_T, (int)N2, NX, &one, r->d_B, (int)N2, r->d_P, 1, &zero, r->d_v, 1);
    if (bs == CUBLAS_STATUS_SUCCESS)
      bs = cublasDgemv(r->blas, CUBLAS_OP_N, (int)N2, NX, &one, r->d_B, (int)N2, r->d_v, 1, &zero, dJ, 1);
    if (bs != CUBLAS_STATUS_SUCCESS) return blas_fail(bs, "cublasDgemv(J)");
  }
  if (K) {
    // K[mu,la] = sum_Q sum_{nu,si} B[mu,nu,Q] P[nu,si] B[la,si,Q]   (computeMatrixExchangeRIK without the N^4 tensor)
    for (int q0 = 0; q0 < NX; q0 += r->qchunk) {
      const int nq = std::min(r->qchunk, NX - q0);
      const double *Bq = r->d_B + N2 * (size_t)q0;
      bs = cublasDgemmStridedBatched(r->blas, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &one, Bq, N, (long long)N2, r->d_P, N, 0,
                                     &zero, r->d_X, N, (long long)N2, nq);
      if (bs != CUBLAS_STATUS_SUCCESS) return blas_fail(bs, "cublasDgemmStridedBatched(X)");
      // K += X (N x N*nq) * Bq (N x N*nq)^T
      bs = cublasDgemm(r->blas, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N * nq, &one, r->d_X, N, Bq, N, q0 == 0 ? &zero : &one, dK, N);
      if (bs != CUBLAS_STATUS_SUCCESS) return blas_fail(bs, "cublasDgemm(K)");
    }
  }
  cudaEventRecord(ev1, e.stream);
  if (J) QLB_CUDA(cudaMemcpyAsync(J, dJ, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  if (K) QLB_CUDA(cudaMemcpyAsync(K, dK, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  if (ms) {
    float t = 0;
    cudaEventElapsedTime(&t, ev0, ev1);
    *ms = t;
  }
  return QLB_OK;
}
