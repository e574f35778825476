// scf.cu -- the SCF step on the device: everything evaluateSCF touches per iteration stays in HBM.
//
// Replaces, per iteration of HartreeFockModule.jl:104-140,
//   evaluateSCFStep (:73-84)            eigen(Symmetric(F), Symmetric(S)); Cocc = C[:, 1:nocc]; P = Cocc*Cocc'
//   computeEnergyHartreeFock (:14-41)   2 tr(TP) + 2 tr(VP) + 2 tr(JP) - tr(KP)
//   computeMatrixFock (SpecialMatricesModule.jl:284-290)   F = T + V + 2J - K
// and keeps P, J, K, F on the device between qlb_build_jk_device calls, so an iteration moves 4 doubles (the energy
// components) to the host instead of 3 N^2.  The generalised symmetric eigenproblem is cuSOLVER's Dsygvd (a library
// call, like the reference's LAPACK); the density P = Cocc Cocc^T is the hand-written DMMA GEMM of dmma_gemm.cuh and
// the four traces are one fused, deterministic two-stage reduction.
#include <cusolverDn.h>

#include <string>
#include <vector>

#include "dmma_gemm.cuh"
#include "qlb_internal.h"

using namespace qlb;

struct qlb_scf {
  qlb_basis *b = nullptr;
  int N = 0, nocc = 0;
  double *d_S = nullptr, *d_T = nullptr, *d_V = nullptr;  // N^2 each
  double *d_P = nullptr, *d_F = nullptr, *d_JK = nullptr; // N^2, N^2, 2 N^2
  double *d_C = nullptr, *d_Swork = nullptr, *d_eps = nullptr;
  double *d_Pbuilt = nullptr, *d_dP = nullptr, *d_dJK = nullptr;  // incremental (density-difference) builds
  double *d_part = nullptr;                                // 4 * NPART partial sums + 4 totals
  double *h_energy = nullptr;                              // pinned, 4 doubles
  double *d_work = nullptr;
  int lwork = 0;
  int *d_info = nullptr;
  cusolverDnHandle_t solver = nullptr;
  bool have_built = false;
  int builds_since_full = 0;
};

namespace {
constexpr int NPART = 256;

// stage 1: per-block partial sums of T.P, V.P, J.P, K.P (fixed block -> element mapping: deterministic)
__global__ void __launch_bounds__(256) trace4_partial_kernel(size_t n, const double *__restrict__ T, const double *__restrict__ V,
                                                              const double *__restrict__ J, const double *__restrict__ K,
                                                              const double *__restrict__ P, double *part) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double p = P[i];
    s[0] = fma(T[i], p, s[0]);
    s[1] = fma(V[i], p, s[1]);
    s[2] = fma(J[i], p, s[2]);
    s[3] = fma(K[i], p, s[3]);
  }
  __shared__ double sh[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
    part[threadIdx.x * NPART + blockIdx.x] = v;
  }
}
// stage 2: one warp per quantity sums the NPART partials in a fixed order; energies = {2 tr TP, 2 tr VP, 2 tr JP, -tr KP}
__global__ void trace4_final_kernel(const double *__restrict__ part, double *out) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0.0;
  for (int i = lane; i < NPART; i += 32) v += part[k * NPART + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out[k] = (k == 3) ? -v : 2.0 * v;
}

__global__ void axpby_kernel(size_t n, double a, const double *__restrict__ x, double b, const double *__restrict__ y, double *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * x[i] + b * y[i];
}
__global__ void fock3_kernel(size_t n, const double *__restrict__ T, const double *__restrict__ V, const double *__restrict__ J,
                             const double *__restrict__ K, double *F) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) F[i] = T[i] + V[i] + 2.0 * J[i] - K[i];  // SpecialMatricesModule.jl:289
}
__global__ void mirror_lower_kernel(int N, double *A) {  // A[j,i] = A[i,j] for i > j
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < N && j < i) A[j + (size_t)N * i] = A[i + (size_t)N * j];
}

int solver_fail(cusolverStatus_t st, const char *what) {
  set_error(std::string("cuSOLVER error ") + std::to_string((int)st) + " in " + what);
  return QLB_ERR_CUDA;
}
}  // namespace

extern "C" int qlb_scf_destroy(qlb_scf *s) {
  if (!s) return QLB_OK;
  if (s->b && s->b->eng) use_engine(s->b->eng);
  if (engine().ready) cudaSetDevice(engine().device);
  void *ptrs[] = {s->d_S, s->d_T, s->d_V, s->d_P, s->d_F, s->d_JK, s->d_C, s->d_Swork, s->d_eps, s->d_Pbuilt, s->d_dP,
                  s->d_dJK, s->d_part, s->d_work, s->d_info};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (s->h_energy) cudaFreeHost(s->h_energy);
  if (s->solver) cusolverDnDestroy(s->solver);
  delete s;
  return QLB_OK;
}

extern "C" int qlb_scf_create(qlb_basis *b, const double *S, const double *T, const double *V, int nocc, qlb_scf **out) {
  if (b && b->eng) use_engine(b->eng);
  Engine &e = engine();
  if (!e.ready) {
    set_error("qlb_init has not been called");
    return QLB_ERR_STATE;
  }
  if (!b || !S || !T || !V || !out || nocc <= 0 || nocc > b->nbf) {
    set_error("qlb_scf_create: bad argument");
    return QLB_ERR_ARG;
  }
  QLB_CUDA(cudaSetDevice(e.device));
  qlb_scf *s = new qlb_scf();
  s->b = b;
  s->N = b->nbf;
  s->nocc = nocc;
  const size_t N2 = (size_t)s->N * s->N;
  double **bufs[] = {&s->d_S, &s->d_T, &s->d_V, &s->d_P, &s->d_F, &s->d_C, &s->d_Swork, &s->d_Pbuilt, &s->d_dP};
  cudaError_t ce = cudaSuccess;
  for (double **p : bufs)
    if (ce == cudaSuccess) ce = cudaMalloc((void **)p, N2 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->d_JK, 2 * N2 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->d_dJK, 2 * N2 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->d_eps, (size_t)s->N * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->d_part, (4 * NPART + 4) * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&s->d_info, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMallocHost((void **)&s->h_energy, 4 * sizeof(double));
  if (ce != cudaSuccess) {
    qlb_scf_destroy(s);
    return cuda_fail(ce, "cudaMalloc(scf)");
  }
  cudaMemcpyAsync(s->d_S, S, N2 * sizeof(double), cudaMemcpyHostToDevice, e.stream);
  cudaMemcpyAsync(s->d_T, T, N2 * sizeof(double), cudaMemcpyHostToDevice, e.stream);
  cudaMemcpyAsync(s->d_V, V, N2 * sizeof(double), cudaMemcpyHostToDevice, e.stream);
  cudaMemsetAsync(s->d_P, 0, N2 * sizeof(double), e.stream);
  cudaMemsetAsync(s->d_JK, 0, 2 * N2 * sizeof(double), e.stream);
  cusolverStatus_t st = cusolverDnCreate(&s->solver);
  if (st == CUSOLVER_STATUS_SUCCESS) st = cusolverDnSetStream(s->solver, e.stream);
  if (st == CUSOLVER_STATUS_SUCCESS)
    st = cusolverDnDsygvd_bufferSize(s->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, s->N, s->d_C,
                                     s->N, s->d_Swork, s->N, s->d_eps, &s->lwork);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    qlb_scf_destroy(s);
    return solver_fail(st, "cusolverDnDsygvd_bufferSize");
  }
  if ((ce = cudaMalloc((void **)&s->d_work, (size_t)std::max(s->lwork, 1) * sizeof(double))) != cudaSuccess ||
      (ce = cudaStreamSynchronize(e.stream)) != cudaSuccess) {
    qlb_scf_destroy(s);
    return cuda_fail(ce, "qlb_scf_create");
  }
  *out = s;
  return QLB_OK;
}

extern "C" int qlb_scf_set_density(qlb_scf *s, const double *P) {
  if (!s || !P) return QLB_ERR_ARG;
  use_engine(s->b->eng);
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)s->N * s->N;
  QLB_CUDA(cudaMemcpyAsync(s->d_P, P, N2 * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  return QLB_OK;
}

// J, K <- P (one fused, screened pass; sharded + all-reduced when a communicator exists), then the four energy
// components of HartreeFockModule.jl:25-35 for (P, J, K).  incremental != 0: J(P) = J(P_built) + J(P - P_built).
extern "C" int qlb_scf_build(qlb_scf *s, double tau, int incremental, double *energies4, qlb_stats *stats) {
  if (!s) return QLB_ERR_ARG;
  use_engine(s->b->eng);
  Engine &e = engine();
  if (!e.ready) return QLB_ERR_STATE;
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)s->N * s->N;
  const unsigned nb = (unsigned)((N2 + 255) / 256), nb2 = (unsigned)((2 * N2 + 255) / 256);
  const bool inc = incremental && s->have_built && s->builds_since_full < 8;
  const double *dens = s->d_P;
  double *jk = s->d_JK;
  if (inc) {
    axpby_kernel<<<nb, 256, 0, e.stream>>>(N2, 1.0, s->d_P, -1.0, s->d_Pbuilt, s->d_dP);
    dens = s->d_dP;
    jk = s->d_dJK;
  }
  if (int rc = qlb_build_jk_device(s->b, dens, tau, jk, jk + N2, e.rank, e.nranks, e.stream, stats)) return rc;
  if (e.nranks > 1)
    if (int rc = qlb_allreduce_jk_device(s->b, jk, e.stream)) return rc;
  if (int rc = qlb_symmetrize_device(s->b, jk, jk + N2, e.stream)) return rc;
  if (inc) {
    axpby_kernel<<<nb2, 256, 0, e.stream>>>(2 * N2, 1.0, s->d_JK, 1.0, s->d_dJK, s->d_JK);
    s->builds_since_full += 1;
  } else {
    s->builds_since_full = 0;
  }
  QLB_CUDA(cudaMemcpyAsync(s->d_Pbuilt, s->d_P, N2 * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  s->have_built = true;
  trace4_partial_kernel<<<NPART, 256, 0, e.stream>>>(N2, s->d_T, s->d_V, s->d_JK, s->d_JK + N2, s->d_P, s->d_part);
  trace4_final_kernel<<<1, 128, 0, e.stream>>>(s->d_part, s->d_part + 4 * NPART);
  QLB_CUDA(cudaGetLastError());
  if (energies4) {
    QLB_CUDA(cudaMemcpyAsync(s->h_energy, s->d_part + 4 * NPART, 4 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    QLB_CUDA(cudaStreamSynchronize(e.stream));
    for (int k = 0; k < 4; ++k) energies4[k] = s->h_energy[k];
  }
  return QLB_OK;
}

// F = T + V + 2J - K; (eps, C) = eigen(F, S); P = C[:, :nocc] C[:, :nocc]^T   (HartreeFockModule.jl:73-84)
extern "C" int qlb_scf_step(qlb_scf *s) {
  if (!s) return QLB_ERR_ARG;
  use_engine(s->b->eng);
  Engine &e = engine();
  if (!e.ready) return QLB_ERR_STATE;
  QLB_CUDA(cudaSetDevice(e.device));
  const int N = s->N;
  const size_t N2 = (size_t)N * N;
  fock3_kernel<<<(unsigned)((N2 + 255) / 256), 256, 0, e.stream>>>(N2, s->d_T, s->d_V, s->d_JK, s->d_JK + N2, s->d_F);
  QLB_CUDA(cudaMemcpyAsync(s->d_C, s->d_F, N2 * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  QLB_CUDA(cudaMemcpyAsync(s->d_Swork, s->d_S, N2 * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  cusolverStatus_t st = cusolverDnDsygvd(s->solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, N, s->d_C, N,
                                         s->d_Swork, N, s->d_eps, s->d_work, s->lwork, s->d_info);
  if (st != CUSOLVER_STATUS_SUCCESS) return solver_fail(st, "cusolverDnDsygvd");
  // P = Cocc Cocc^T: lower triangle by DMMA tiles, then mirrored (exactly symmetric, as the build assumes)
  cudaError_t ce = dmma::dgemm(e.stream, false, true, N, N, s->nocc, 1.0, s->d_C, N, s->d_C, N, 0.0, s->d_P, N, 1, 0, 0, 0, true);
  if (ce != cudaSuccess) return cuda_fail(ce, "dgemm(density)");
  dim3 block(32, 8), grid((N + 31) / 32, (N + 7) / 8);
  mirror_lower_kernel<<<grid, block, 0, e.stream>>>(N, s->d_P);
  int info = 0;
  QLB_CUDA(cudaMemcpyAsync(&info, s->d_info, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  if (info != 0) {
    set_error("cusolverDnDsygvd: info = " + std::to_string(info) + " (overlap matrix not positive definite, or no convergence)");
    return QLB_ERR_CUDA;
  }
  return QLB_OK;
}

// what: 0 P, 1 F, 2 J, 3 K, 4 orbital energies (N), 5 orbital coefficients C (N x N)
extern "C" int qlb_scf_get(qlb_scf *s, int what, double *out) {
  if (!s || !out) return QLB_ERR_ARG;
  use_engine(s->b->eng);
  Engine &e = engine();
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)s->N * s->N;
  const double *src = nullptr;
  size_t n = N2;
  switch (what) {
    case 0: src = s->d_P; break;
    case 1: src = s->d_F; break;
    case 2: src = s->d_JK; break;
    case 3: src = s->d_JK + N2; break;
    case 4: src = s->d_eps; n = s->N; break;
    case 5: src = s->d_C; break;
    default: set_error("qlb_scf_get: unknown item"); return QLB_ERR_ARG;
  }
  QLB_CUDA(cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  return QLB_OK;
}

// C = alpha op(A) op(B) + beta C on DEVICE pointers with the library's DMMA kernel (tests, benchmarks)
extern "C" int qlb_dgemm_device(int ta, int tb, int M, int N, int K, double alpha, const double *dA, int lda, const double *dB, int ldb,
                                double beta, double *dC, int ldc, void *stream_) {
  Engine &e = engine();
  if (!e.ready) return QLB_ERR_STATE;
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : e.stream;
  cudaError_t ce = dmma::dgemm(st, ta != 0, tb != 0, M, N, K, alpha, dA, lda, dB, ldb, beta, dC, ldc);
  if (ce != cudaSuccess) return cuda_fail(ce, "dgemm");
  return QLB_OK;
}

// The two packed-symmetric products of the RI-K contraction on DEVICE pointers (tests): Bp = nslab packed lower triangles
// (column-major, N(N+1)/2 doubles each).
//   side 0:  X[:, :, q] = B^q R            R: N x ncol,          X: N x ncol x nslab
//   side 1:  Kout = sum_q X[:, :, q] B^q   X: N x N x nslab,     Kout: N x N
extern "C" int qlb_dsymm_packed_device(int side, int N, int ncol, int nslab, const double *dBp, const double *dR, double *dOut, void *stream_) {
  Engine &e = engine();
  if (!e.ready) return QLB_ERR_STATE;
  if (N <= 0 || nslab <= 0 || !dBp || !dR || !dOut) {
    set_error("qlb_dsymm_packed_device: bad argument");
    return QLB_ERR_ARG;
  }
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : e.stream;
  const int64_t NP = (int64_t)N * (N + 1) / 2;
  cudaError_t ce;
  if (side == 0)
    ce = dmma::dgemm_ex(st, N, ncol, N, 1.0, {dBp, NP, dmma::LAY_SYM, NP}, {dR, (int64_t)N, dmma::LAY_K, 0}, 0.0, dOut, (int64_t)N,
                        (int64_t)N * ncol, nslab, false, N);
  else
    ce = dmma::dgemm_ex(st, N, N, N * nslab, 1.0, {dR, (int64_t)N, dmma::LAY_MN, 0}, {dBp, NP, dmma::LAY_SYM, 0}, 0.0, dOut, (int64_t)N, 0, 1,
                        false, N);
  if (ce != cudaSuccess) return cuda_fail(ce, "dsymm_packed");
  return QLB_OK;
}

// DMMA microbenchmark: the measured FP64 tensor-core roofline denominator of the RI contractions.  8 independent
// accumulator pairs per warp, operands in registers, no memory traffic.
namespace {
__global__ void dmma_probe_kernel(double *out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma::mma_m8n8k4(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int qlb_fp64_tensor_peak_probe(double *tflops) {
  Engine &e = engine();
  if (!e.ready) return QLB_ERR_STATE;
  QLB_CUDA(cudaSetDevice(e.device));
  const int blocks = e.sm_count * 8, threads = 256, iters = 1 << 13;
  double *d = nullptr;
  QLB_CUDA(cudaMalloc((void **)&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  dmma_probe_kernel<<<blocks, threads, 0, e.stream>>>(d, 256);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, e.stream);
    dmma_probe_kernel<<<blocks, threads, 0, e.stream>>>(d, iters);
    cudaEventRecord(b, e.stream);
    QLB_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    // one m8n8k4 = 8*8*4 FMAs = 512 flops per warp
    const double fl = 512.0 * 8.0 * (double)iters * blocks * (threads / 32);
    best = std::max(best, fl / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  if (tflops) *tflops = best;
  return QLB_OK;
}
