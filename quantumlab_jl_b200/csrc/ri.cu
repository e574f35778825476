// ri.cu -- resolution of the identity in the Coulomb metric.
//
// Replaces ResolutionIdentityModule.jl: computeMatrixResolutionOfTheIdentityCoulombMetric (:41-50),
// computeTensorElectronRepulsionIntegralsRICoulomb (:54-58) and computeMatrixExchangeRIK (:69-78).  The reference
// forms (mu nu|P) and (P|Q) through its 4-centre ERI with the unit s function `identityCGF`
// (BasisFunctionsModule.jl:50-51), takes B = (mu nu|P) (P|Q)^-1/2, rebuilds the FULL N^4 tensor B B^T and contracts it.
// Here:
//   * the 3- and 2-index integrals come from the class kernels (MODE 2, an (auxiliary shell, unit s) pair on one side),
//     auxiliary shells up to f;
//   * B is stored ONCE, packed over mu >= nu: B[pk(mu,nu) + NP*Q], NP = N(N+1)/2.  It is accumulated chunk of auxiliary
//     shells by chunk, B += (mu nu|P_chunk) (P|Q)^-1/2 [chunk, :], so the raw 3-index array only ever exists as a
//     chunk-sized scratch;
//   * every dense contraction is the library's own DMMA kernel (dmma_gemm.cuh; LAY_SYM reads the packed B^Q as a
//     symmetric N x N operand), the two matrix-vector products of RI-J are bandwidth-bound streaming kernels:
//         RI-J   v_Q = sum_{mu>=nu} B[mu nu,Q] w P[mu nu];   J[mu nu] = sum_Q B[mu nu,Q] v_Q          (2 passes over B: HBM)
//         RI-K   general P:   X^Q = B^Q P,  K = sum_Q X^Q B^Q                                    (2 N^3 Naux x 2 flops)
//                occupied C:  X^Q = B^Q C_occ,  K = sum_Q X^Q X^Q^T                              (3 N^2 nocc Naux flops)
//     the N^4 tensor is only formed by qlb_ri_eri_tensor (N <= 64, tests);
//   * with a communicator every rank owns a contiguous slice of the auxiliary index Q (its columns of B); J and K are
//     sums over Q, so the ranks' partial [J;K] are combined by ONE all-reduce, exactly like the conventional build.
// Only the eigen-decomposition of the (small) metric is a library call (cuSOLVER Dsyevd).
#include <cusolverDn.h>

#include <algorithm>
#include <string>
#include <vector>

#include "dmma_gemm.cuh"
#include "qlb_internal.h"

using namespace qlb;

struct qlb_ri {
  qlb_basis *comb = nullptr;  // orbital shells + auxiliary shells + unit s function
  int N = 0, Naux = 0, aux_fn0 = 0;
  int64_t NP = 0;             // N (N+1) / 2
  int q0 = 0, nq = 0;         // this rank's slice of the auxiliary functions
  double *d_B = nullptr;      // NP x nq, packed over mu >= nu
  double *d_P = nullptr, *d_pw = nullptr, *d_Jp = nullptr, *d_JK = nullptr, *d_v = nullptr, *d_X = nullptr, *d_C = nullptr;
  int qchunk = 0;
  cusolverDnHandle_t solver = nullptr;
};

namespace {

__global__ void scale_columns_kernel(int n, const double *__restrict__ w, double *V) {  // V[:, j] *= w[j]
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n) V[i + (size_t)n * j] *= w[j];
}

// pw[pk(a,b)] = w P[a,b], w = 1 on the diagonal, 2 off it (the packed sum runs over a >= b only)
__global__ void pack_weighted_kernel(int N, const double *__restrict__ P, double *pw) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a < N && b <= a) pw[dmma::sym_pk(a, b, N)] = (a == b) ? P[a + (size_t)N * a] : P[a + (size_t)N * b] + P[b + (size_t)N * a];
}
__global__ void unpack_sym_kernel(int N, const double *__restrict__ p, double *A) {  // A[a,b] = A[b,a] = p[pk(a,b)]
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a < N && b <= a) {
    const double v = p[dmma::sym_pk(a, b, N)];
    A[a + (size_t)N * b] = v;
    A[b + (size_t)N * a] = v;
  }
}
__global__ void mirror_lower_kernel(int N, double *A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < N && j < i) A[j + (size_t)N * i] = A[i + (size_t)N * j];
}
// v[q] = sum_pk B[pk + NP q] pw[pk]: one CTA per auxiliary function streams its column (coalesced)
__global__ void __launch_bounds__(256) gemv_t_kernel(int64_t NP, const double *__restrict__ B, const double *__restrict__ pw, double *v) {
  const double *col = B + NP * blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < NP; i += 256) s = fma(col[i], pw[i], s);
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    v[blockIdx.x] = t;
  }
}
// out[pk] = sum_q B[pk + NP q] v[q]: one thread per packed pair, the warp's loads of every column are coalesced
__global__ void __launch_bounds__(256) gemv_n_kernel(int64_t NP, int nq, const double *__restrict__ B, const double *__restrict__ v, double *out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= NP) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int q = 0;
  for (; q + 4 <= nq; q += 4) {
    s0 = fma(B[i + NP * q], v[q], s0);
    s1 = fma(B[i + NP * (q + 1)], v[q + 1], s1);
    s2 = fma(B[i + NP * (q + 2)], v[q + 2], s2);
    s3 = fma(B[i + NP * (q + 3)], v[q + 3], s3);
  }
  for (; q < nq; ++q) s0 = fma(B[i + NP * q], v[q], s0);
  out[i] = (s0 + s1) + (s2 + s3);
}
// dense[a + N b + N^2 q] = B[pk(a,b) + NP q]  (tests: qlb_ri_b_tensor / qlb_ri_eri_tensor)
__global__ void unpack_b_kernel(int N, int64_t NP, const double *__restrict__ B, double *dense) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, q = blockIdx.z;
  if (a < N) dense[a + (size_t)N * b + (size_t)N * N * q] = B[dmma::sym_pk(a, b, N) + NP * q];
}

int solver_fail(cusolverStatus_t st, const char *what) {
  set_error(std::string("cuSOLVER error ") + std::to_string((int)st) + " in " + what);
  return QLB_ERR_CUDA;
}

// launch MODE-2 class kernels: bra pair range [xlo, xhi) of class X vs ket pair range [ylo, yhi) of class Y
int ri_launch(ClassParams &prm, int X, int xlo, int xhi, int Y, int ylo, int yhi, int ri_mode, cudaStream_t s) {
  prm.bra_off = xlo; prm.bra_n = xhi - xlo;
  prm.ket_off = ylo; prm.ket_n = yhi - ylo;
  if (prm.bra_n <= 0 || prm.ket_n <= 0) return QLB_OK;
  prm.same = 0;
  prm.screen = 0;
  prm.ri_mode = ri_mode;
  prm.nbx = (prm.ket_n + TILE_KET - 1) / TILE_KET;
  prm.ysplit = prm.nbx;
  prm.rank = 0;
  prm.nranks = 1;
  const int64_t grid = (int64_t)((prm.bra_n + TILE_BRA - 1) / TILE_BRA) * prm.nbx;
  if (int ce = launch_class_kernel(X, Y, 2, (unsigned)grid, s, prm)) return cuda_fail((cudaError_t)ce, "eri_class_kernel(ri)");
  return QLB_OK;
}

}  // namespace

extern "C" int qlb_ri_destroy(qlb_ri *r) {
  if (!r) return QLB_OK;
  if (r->comb && r->comb->eng) use_engine(r->comb->eng);
  if (engine().ready) cudaSetDevice(engine().device);
  void *ptrs[] = {r->d_B, r->d_P, r->d_pw, r->d_Jp, r->d_JK, r->d_v, r->d_X, r->d_C};
  for (void *p : ptrs)
    if (p) cudaFree(p);
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

extern "C" int qlb_ri_slice(const qlb_ri *r, int *q0, int *nq) {
  if (!r || !q0 || !nq) return QLB_ERR_ARG;
  *q0 = r->q0;
  *nq = r->nq;
  return QLB_OK;
}

extern "C" int qlb_ri_create(qlb_basis *orb, int naux_sh, const double *centers, const int32_t *L, const int32_t *nprim,
                             const double *exps, const double *coefs, qlb_ri **out) {
  if (orb && orb->eng) use_engine(orb->eng);
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
  r->NP = (int64_t)r->N * (r->N + 1) / 2;
  r->aux_fn0 = b->h_boff[orb->nsh];
  r->Naux = b->h_boff[nsh - 1] - r->aux_fn0;
  r->q0 = (int)((int64_t)r->Naux * e.rank / e.nranks);
  r->nq = (int)((int64_t)r->Naux * (e.rank + 1) / e.nranks) - r->q0;
  const size_t N = r->N, NX = r->Naux, N2 = N * N, NP = (size_t)r->NP;
  cudaError_t ce;
  double *d_S = nullptr, *d_M = nullptr, *d_w = nullptr, *d_Mh = nullptr, *d_work = nullptr;
  int *d_info = nullptr;
  auto cleanup = [&]() {
    void *ptrs[] = {d_S, d_M, d_w, d_Mh, d_work, d_info};
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
  // auxiliary shells are processed in chunks of about `cfun` functions: the raw (mu nu|P) only exists as NP x chunk scratch
  const size_t scratch_cap = (size_t)2 << 30;  // bytes
  const int cfun = (int)std::max<size_t>(16, std::min<size_t>(NX, scratch_cap / (NP * sizeof(double))));
  int max_chunk_fn = 0;
  std::vector<std::pair<int, int>> chunks;  // [first aux shell, one past last) relative to the first auxiliary shell
  for (int s0 = 0; s0 < naux_sh;) {
    int s1 = s0, nf = 0;
    while (s1 < naux_sh && (nf == 0 || nf + ncart(L[s1]) <= cfun)) nf += ncart(L[s1++]);
    chunks.push_back({s0, s1});
    max_chunk_fn = std::max(max_chunk_fn, nf);
    s0 = s1;
  }
  RI_CUDA(cudaMalloc((void **)&r->d_B, NP * std::max<size_t>(r->nq, 1) * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_S, NP * (size_t)max_chunk_fn * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_M, NX * NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_Mh, NX * NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_w, NX * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&d_info, sizeof(int)));
  RI_CUDA(cudaMalloc((void **)&r->d_P, N2 * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_C, N2 * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_pw, NP * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_Jp, NP * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_JK, (2 * N2 + QLB_JK_TAIL) * sizeof(double)));
  RI_CUDA(cudaMalloc((void **)&r->d_v, std::max<size_t>(r->nq, 1) * sizeof(double)));
  // chunk of auxiliary functions for the exchange contraction: X^Q for qchunk Q's at a time (<= 1 GiB scratch)
  r->qchunk = (int)std::max<size_t>(1, std::min<size_t>(std::max(r->nq, 1), ((size_t)1 << 30) / (N2 * sizeof(double))));
  RI_CUDA(cudaMalloc((void **)&r->d_X, N2 * r->qchunk * sizeof(double)));
  RI_CUDA(cudaMemsetAsync(d_M, 0, NX * NX * sizeof(double), e.stream));
  ClassParams prm;
  base_class_params(b, prm);
  prm.ri_n = (int)N; prm.ri_naux = (int)NX; prm.ri_aux_fn0 = r->aux_fn0;
  // ---- 2-index (P|Q) integrals: every rank needs the whole (small) metric
  for (int Y1 = 0; Y1 < NCLASS_AUX && rc == QLB_OK; ++Y1)
    for (int Y2 = 0; Y2 <= Y1 && rc == QLB_OK; ++Y2) {
      prm.tensor = d_M;
      rc = ri_launch(prm, Y1, b->ri_class_off[Y1], b->ri_class_off[Y1 + 1], Y2, b->ri_class_off[Y2], b->ri_class_off[Y2 + 1], 2, e.stream);
    }
  if (rc) {
    cleanup();
    qlb_ri_destroy(r);
    return rc;
  }
  // ---- (P|Q)^-1/2 = V diag(lambda^-1/2) V^T  (the reference: sqrt(inv(C)), ResolutionIdentityModule.jl:49)
  cusolverStatus_t ss = cusolverDnCreate(&r->solver);
  if (ss != CUSOLVER_STATUS_SUCCESS) {
    cleanup();
    qlb_ri_destroy(r);
    return solver_fail(ss, "cusolverDnCreate");
  }
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
  RI_CUDA(dmma::dgemm(e.stream, false, true, (int)NX, (int)NX, (int)NX, 1.0, d_M, (int64_t)NX, d_M, (int64_t)NX, 0.0, d_Mh, (int64_t)NX));
  // ---- B[:, slice] = sum over chunks of (mu nu|P_chunk) M^-1/2[chunk, slice]
  const int n_orb = orb->nsh;
  bool first = true;
  for (const auto &ch : chunks) {
    const int sh0 = n_orb + ch.first, sh1 = n_orb + ch.second;  // auxiliary shells of this chunk (indices in the combined list)
    const int f0 = b->h_boff[sh0] - r->aux_fn0, f1 = b->h_boff[sh1] - r->aux_fn0;
    RI_CUDA(cudaMemsetAsync(d_S, 0, NP * (size_t)(f1 - f0) * sizeof(double), e.stream));
    prm.tensor = d_S;
    prm.ri_aux_fn0 = r->aux_fn0 + f0;
    for (int Y = 0; Y < NCLASS_AUX && rc == QLB_OK; ++Y) {
      // (auxiliary, unit) pairs of class Y are in shell order: the chunk's pairs are a contiguous sub-range
      const int *pa = b->h_pair_a.data();
      const int ylo = (int)(std::lower_bound(pa + b->ri_class_off[Y], pa + b->ri_class_off[Y + 1], sh0) - pa);
      const int yhi = (int)(std::lower_bound(pa + b->ri_class_off[Y], pa + b->ri_class_off[Y + 1], sh1) - pa);
      for (int X = 0; X < NCLASS && rc == QLB_OK; ++X) {
        // X: orbital pair class, Y: (aux, unit) pair class.  Kernels exist for bra class >= ket class.
        if (X >= Y) rc = ri_launch(prm, X, b->class_off[X], b->class_off[X + 1], Y, ylo, yhi, 0, e.stream);
        else rc = ri_launch(prm, Y, ylo, yhi, X, b->class_off[X], b->class_off[X + 1], 1, e.stream);
      }
    }
    if (rc) {
      cleanup();
      qlb_ri_destroy(r);
      return rc;
    }
    if (r->nq > 0)
      RI_CUDA(dmma::dgemm(e.stream, false, false, (int)NP, r->nq, f1 - f0, 1.0, d_S, (int64_t)NP, d_Mh + f0 + NX * (size_t)r->q0, (int64_t)NX,
                          first ? 0.0 : 1.0, r->d_B, (int64_t)NP));
    first = false;
  }
  RI_CUDA(cudaStreamSynchronize(e.stream));
  cleanup();
#undef RI_CUDA
  *out = r;
  return QLB_OK;
}

// B as the reference's N x N x Naux array; with a communicator only this rank's slice of Q is written.
extern "C" int qlb_ri_b_tensor(qlb_ri *r, double *B) {
  if (!r || !B) return QLB_ERR_ARG;
  use_engine(r->comb->eng);
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)r->N * r->N;
  if (r->nq == 0) return QLB_OK;
  double *dense = nullptr;
  QLB_CUDA(cudaMalloc((void **)&dense, N2 * r->nq * sizeof(double)));
  unpack_b_kernel<<<dim3((r->N + 63) / 64, r->N, r->nq), 64, 0, e.stream>>>(r->N, r->NP, r->d_B, dense);
  cudaError_t ce = cudaMemcpyAsync(B + N2 * r->q0, dense, N2 * r->nq * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
  cudaFree(dense);
  if (ce != cudaSuccess) return cuda_fail(ce, "qlb_ri_b_tensor");
  return QLB_OK;
}

// ERI_RI[mu,nu,la,si] = sum_Q B[mu,nu,Q] B[la,si,Q]  (computeTensorElectronRepulsionIntegralsRICoulomb, :54-58; tests only)
extern "C" int qlb_ri_eri_tensor(qlb_ri *r, double *out) {
  if (!r || !out) return QLB_ERR_ARG;
  if (r->N > 64) {
    set_error("qlb_ri_eri_tensor: N > 64; the N^4 tensor is only materialised for test-sized systems");
    return QLB_ERR_ARG;
  }
  use_engine(r->comb->eng);
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)r->N * r->N;
  double *dT = nullptr, *dense = nullptr;
  QLB_CUDA(cudaMalloc((void **)&dT, N2 * N2 * sizeof(double)));
  cudaError_t ce = cudaMalloc((void **)&dense, N2 * std::max(r->nq, 1) * sizeof(double));
  if (ce == cudaSuccess) {
    unpack_b_kernel<<<dim3((r->N + 63) / 64, r->N, std::max(r->nq, 1)), 64, 0, e.stream>>>(r->N, r->NP, r->d_B, dense);
    ce = dmma::dgemm(e.stream, false, true, (int)N2, (int)N2, r->nq, 1.0, dense, (int64_t)N2, dense, (int64_t)N2, 0.0, dT, (int64_t)N2);
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, dT, N2 * N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
  cudaFree(dT);
  if (dense) cudaFree(dense);
  if (ce != cudaSuccess) return cuda_fail(ce, "qlb_ri_eri_tensor");
  return QLB_OK;
}

// shared tail of the two build functions: J from d_P, K from either the general-P or the occupied-orbital form
static int ri_build_common(qlb_ri *r, int nocc, bool want_j, bool want_k, double *J, double *K, double *ms) {
  Engine &e = engine();
  const int N = r->N, nq = r->nq;
  const int64_t NP = r->NP;
  const size_t N2 = (size_t)N * N;
  cudaEvent_t ev0 = r->comb->ev[0], ev1 = r->comb->ev[2];
  cudaEventRecord(ev0, e.stream);
  double *dJ = r->d_JK, *dK = r->d_JK + N2;
  const dim3 pgrid((N + 127) / 128, N);
  if (want_j) {
    // v_Q = sum_{la>=si} B[la si, Q] w P[la si];  J[mu nu] = sum_Q B[mu nu, Q] v_Q
    pack_weighted_kernel<<<pgrid, 128, 0, e.stream>>>(N, r->d_P, r->d_pw);
    if (nq > 0) {
      gemv_t_kernel<<<nq, 256, 0, e.stream>>>(NP, r->d_B, r->d_pw, r->d_v);
      gemv_n_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, e.stream>>>(NP, nq, r->d_B, r->d_v, r->d_Jp);
    } else {
      QLB_CUDA(cudaMemsetAsync(r->d_Jp, 0, NP * sizeof(double), e.stream));
    }
    unpack_sym_kernel<<<pgrid, 128, 0, e.stream>>>(N, r->d_Jp, dJ);
  } else {
    QLB_CUDA(cudaMemsetAsync(dJ, 0, N2 * sizeof(double), e.stream));
  }
  if (want_k && nq > 0) {
    const int ncol = nocc > 0 ? nocc : N;  // columns of X^Q
    const double *rhs = nocc > 0 ? r->d_C : r->d_P;
    for (int q0 = 0; q0 < nq; q0 += r->qchunk) {
      const int nqc = std::min(r->qchunk, nq - q0);
      const double *Bq = r->d_B + NP * (size_t)q0;
      // X^Q = B^Q rhs for the chunk's Q's (batched; B^Q read packed as a symmetric operand)
      cudaError_t ce = dmma::dgemm_ex(e.stream, N, ncol, N, 1.0, {Bq, NP, dmma::LAY_SYM, NP}, {rhs, (int64_t)N, dmma::LAY_K, 0}, 0.0, r->d_X,
                                      (int64_t)N, (int64_t)N * ncol, nqc, false, N);
      if (ce != cudaSuccess) return cuda_fail(ce, "dgemm(X = B^Q rhs)");
      if (nocc > 0) {
        // K += X X^T over (i, Q): lower triangle only, mirrored below
        ce = dmma::dgemm(e.stream, false, true, N, N, ncol * nqc, 1.0, r->d_X, (int64_t)N, r->d_X, (int64_t)N, q0 == 0 ? 0.0 : 1.0, dK,
                         (int64_t)N, 1, 0, 0, 0, true);
      } else {
        // K[mu,la] += sum_{si,Q} X^Q[mu,si] B^Q[la,si]
        ce = dmma::dgemm_ex(e.stream, N, N, N * nqc, 1.0, {r->d_X, (int64_t)N, dmma::LAY_MN, 0}, {Bq, NP, dmma::LAY_SYM, 0}, q0 == 0 ? 0.0 : 1.0,
                            dK, (int64_t)N, 0, 1, false, N);
      }
      if (ce != cudaSuccess) return cuda_fail(ce, "dgemm(K)");
    }
    if (nocc > 0) mirror_lower_kernel<<<dim3((N + 31) / 32, (N + 7) / 8), dim3(32, 8), 0, e.stream>>>(N, dK);
  } else {
    QLB_CUDA(cudaMemsetAsync(dK, 0, N2 * sizeof(double), e.stream));
  }
  if (e.nranks > 1)  // the ranks' slices of Q: one all-reduce of [J;K], as in the conventional build
    if (int rc = allreduce_sum_device(r->d_JK, 2 * N2, e.stream)) return rc;
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

extern "C" int qlb_ri_build_jk(qlb_ri *r, const double *P, double *J, double *K, double *ms) {
  if (!r || !P || (!J && !K)) {
    set_error("qlb_ri_build_jk: null pointer");
    return QLB_ERR_ARG;
  }
  use_engine(r->comb->eng);
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  QLB_CUDA(cudaMemcpyAsync(r->d_P, P, (size_t)r->N * r->N * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  return ri_build_common(r, 0, J != nullptr, K != nullptr, J, K, ms);
}

// occupied-orbital form: P = C_occ C_occ^T (built on the device); RI-K costs 3 N^2 nocc Naux flops instead of 4 N^3 Naux
extern "C" int qlb_ri_build_jk_occ(qlb_ri *r, const double *Cocc, int nocc, double *J, double *K, double *ms) {
  if (!r || !Cocc || nocc <= 0 || nocc > r->N || (!J && !K)) {
    set_error("qlb_ri_build_jk_occ: bad argument");
    return QLB_ERR_ARG;
  }
  use_engine(r->comb->eng);
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const int N = r->N;
  QLB_CUDA(cudaMemcpyAsync(r->d_C, Cocc, (size_t)N * nocc * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  cudaError_t ce = dmma::dgemm(e.stream, false, true, N, N, nocc, 1.0, r->d_C, (int64_t)N, r->d_C, (int64_t)N, 0.0, r->d_P, (int64_t)N);
  if (ce != cudaSuccess) return cuda_fail(ce, "dgemm(P = C C^T)");
  // X^Q needs N * nocc * qchunk doubles: the scratch was sized for N * N * qchunk
  return ri_build_common(r, nocc, J != nullptr, K != nullptr, J, K, ms);
}
