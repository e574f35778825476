// dmma_gemm.cuh -- hand-written FP64 tensor-core GEMM (mma.sync.aligned.m8n8k4 ... f64 = DMMA in SASS) for the genuinely
// dense contractions of the path:
//   P = C_occ C_occ^T                       (evaluateSCFStep, HartreeFockModule.jl:82)
//   RI-J / RI-K 3-index contractions        (ResolutionIdentityModule.jl:41-78)
// C[M x N] = alpha * op(A)[M x K] * op(B)[K x N] + beta * C, everything column-major FP64.
// CTA tile 64 x 64, K step 16, 4 warps (2 x 2), each warp a 32 x 32 block = 4 x 4 DMMA tiles (32 accumulator doubles
// per lane); operands are staged through shared memory with cp.async (two stages), laid out so that the global reads
// are contiguous whichever operand is transposed and the fragment reads hit the 2-wavefront minimum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb {
namespace dmma {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int LDM = BM + 8;  // [k][m] layout: row stride 72 doubles (k rows land 16 banks apart)
constexpr int LDK = BK + 4;  // [m][k] layout: row stride 20 doubles
constexpr int STAGE_DOUBLES = (BK * LDM > BM * LDK) ? BK * LDM : BM * LDK;

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src, bool pred) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 8 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void mma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// one operand tile (64 x 16 of op(X)) into shared memory.  CONTIG_MN: the operand's memory is contiguous along its
// M (or N) index -> layout [k][m]; else contiguous along k -> layout [m][k].
template <bool CONTIG_MN>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ g, int64_t ld, int mn0, int k0, int MN, int K, int tid) {
  if (CONTIG_MN) {
    // element (mn, k) at g[mn + ld*k]
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + 128 * i, mn = e % BM, k = e / BM;
      const bool ok = (mn0 + mn < MN) && (k0 + k < K);
      cp_async8(s + k * LDM + mn, g + (ok ? (int64_t)(mn0 + mn) + ld * (k0 + k) : 0), ok);
    }
  } else {
    // element (mn, k) at g[k + ld*mn]
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + 128 * i, k = e % BK, mn = e / BK;
      const bool ok = (mn0 + mn < MN) && (k0 + k < K);
      cp_async8(s + mn * LDK + k, g + (ok ? (int64_t)(k0 + k) + ld * (mn0 + mn) : 0), ok);
    }
  }
}
template <bool CONTIG_MN>
__device__ __forceinline__ double frag(const double *s, int mn, int k) {
  return CONTIG_MN ? s[k * LDM + mn] : s[mn * LDK + k];
}

// TA: op(A) = A^T (A stored K x M, lda >= K); TB: op(B) = B^T (B stored N x K, ldb >= N).
// gridDim.z batches: operand pointers advance by strideA/strideB/strideC doubles per batch.
// LOWER_ONLY: skip CTA tiles strictly above the diagonal (symmetric results; the caller mirrors).
template <bool TA, bool TB, bool LOWER_ONLY>
__global__ void __launch_bounds__(128) dgemm_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int64_t lda,
                                                    int64_t strideA, const double *__restrict__ B, int64_t ldb, int64_t strideB,
                                                    double beta, double *C, int64_t ldc, int64_t strideC) {
  __shared__ double As[2][STAGE_DOUBLES], Bs[2][STAGE_DOUBLES];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (LOWER_ONLY && n0 > m0 + BM - 1) return;
  A += strideA * blockIdx.z;
  B += strideB * blockIdx.z;
  C += strideC * blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int fr = lane >> 2, fc = lane & 3;  // fragment row / column
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nk = (K + BK - 1) / BK;
  // A is contiguous along m unless transposed; B (K x N) is contiguous along k unless transposed
  load_tile<!TA>(As[0], A, lda, m0, 0, M, K, tid);
  load_tile<TB>(Bs[0], B, ldb, n0, 0, N, K, tid);
  cp_async_commit();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<!TA>(As[cur ^ 1], A, lda, m0, (kt + 1) * BK, M, K, tid);
      load_tile<TB>(Bs[cur ^ 1], B, ldb, n0, (kt + 1) * BK, N, K, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = frag<!TA>(As[cur], wm + 8 * i + fr, kk + fc);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = frag<TB>(Bs[cur], wn + 8 * j + fr, kk + fc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // C fragment: lane holds (row fr, columns 2*fc, 2*fc+1) of each 8 x 8 tile
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = m0 + wm + 8 * i + fr, n = n0 + wn + 8 * j + 2 * fc + h;
        if (m < M && n < N) {
          double *c = C + (int64_t)m + ldc * n;
          *c = (beta == 0.0) ? alpha * acc[i][j][h] : alpha * acc[i][j][h] + beta * *c;
        }
      }
}

// host launcher.  Returns the cudaError_t of the launch.
inline cudaError_t dgemm(cudaStream_t s, bool ta, bool tb, int M, int N, int K, double alpha, const double *A, int64_t lda,
                         const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int batch = 1, int64_t strideA = 0,
                         int64_t strideB = 0, int64_t strideC = 0, bool lower_only = false) {
  if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
  const dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, batch), block(128);
#define QLB_DGEMM_CASE(TA, TB)                                                                                                   \
  if (ta == TA && tb == TB) {                                                                                                    \
    if (lower_only) dgemm_kernel<TA, TB, true><<<grid, block, 0, s>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC); \
    else dgemm_kernel<TA, TB, false><<<grid, block, 0, s>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC);           \
  }
  QLB_DGEMM_CASE(false, false)
  QLB_DGEMM_CASE(false, true)
  QLB_DGEMM_CASE(true, false)
  QLB_DGEMM_CASE(true, true)
#undef QLB_DGEMM_CASE
  return cudaGetLastError();
}

}  // namespace dmma
}  // namespace qlb
