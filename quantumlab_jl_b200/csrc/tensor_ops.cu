// tensor_ops.cu -- contraction of a STORED N^4 ERI tensor with the density on the device: the "tensor flavour" of
// computeMatrixCoulomb / computeMatrixExchange (SpecialMatricesModule.jl:227-229, 264-266: one @tensor contraction of
// the N^4 array, 8 N^4 bytes read per call -- HBM-bound).  Test-sized N only (the engine itself never forms N^4).
//   J[m,n] = sum_{l,s} T[m,n,l,s] P[l,s]          K[m,n] = sum_{l,s} T[m,l,n,s] P[l,s]
// One thread per output row index m (the fastest tensor index: coalesced reads), one CTA per (n, chunk of (l,s)).
#include "qlb_internal.h"

using namespace qlb;

namespace {
template <bool EXCHANGE>
__global__ void tensor_contract_kernel(int N, const double *__restrict__ T, const double *__restrict__ P, double *out, int chunk) {
  const int n = blockIdx.x;
  const size_t N1 = N, N2 = N1 * N1, N3 = N2 * N1;
  const int ls0 = blockIdx.y * chunk, ls1 = min(ls0 + chunk, N * N);
  for (int m = threadIdx.x; m < N; m += blockDim.x) {
    double acc = 0.0;
    for (int ls = ls0; ls < ls1; ++ls) {
      const int l = ls % N, s = ls / N;
      const double t = EXCHANGE ? T[m + N1 * l + N2 * n + N3 * s] : T[m + N1 * n + N2 * l + N3 * s];
      acc = fma(t, P[l + N1 * s], acc);
    }
    atomicAdd(out + m + N1 * n, acc);
  }
}
}  // namespace

extern "C" int qlb_tensor_contract(int N, const double *T, const double *P, int exchange, double *out) {
  Engine &e = engine();
  if (!e.ready) {
    set_error("qlb_init has not been called (no CPU fallback)");
    return QLB_ERR_STATE;
  }
  if (N <= 0 || N > 160 || !T || !P || !out) {
    set_error("qlb_tensor_contract: bad argument (N <= 160: the N^4 tensor is a test-sized object)");
    return QLB_ERR_ARG;
  }
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t n2 = (size_t)N * N, n4 = n2 * n2;
  double *dT = nullptr, *dP = nullptr, *dO = nullptr;
  cudaError_t ce = cudaMalloc((void **)&dT, n4 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&dP, n2 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc((void **)&dO, n2 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(dT, T, n4 * sizeof(double), cudaMemcpyHostToDevice, e.stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(dP, P, n2 * sizeof(double), cudaMemcpyHostToDevice, e.stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(dO, 0, n2 * sizeof(double), e.stream);
  if (ce == cudaSuccess) {
    const int chunk = 256;
    const dim3 grid(N, (unsigned)((n2 + chunk - 1) / chunk));
    if (exchange) tensor_contract_kernel<true><<<grid, 128, 0, e.stream>>>(N, dT, dP, dO, chunk);
    else tensor_contract_kernel<false><<<grid, 128, 0, e.stream>>>(N, dT, dP, dO, chunk);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, dO, n2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
  cudaFree(dT);
  cudaFree(dP);
  cudaFree(dO);
  if (ce != cudaSuccess) return cuda_fail(ce, "qlb_tensor_contract");
  return QLB_OK;
}
