// dmma_gemm.cuh -- hand-written FP64 tensor-core GEMM (mma.sync.aligned.m8n8k4 ... f64 = DMMA in SASS) for the genuinely
// dense contractions of the path:
//   P = C_occ C_occ^T                       (evaluateSCFStep, HartreeFockModule.jl:82)
//   RI-J / RI-K 3-index contractions        (ResolutionIdentityModule.jl:41-78)
// C[M x N] = alpha * op(A)[M x K] * op(B)[K x N] + beta * C, everything column-major FP64.
// CTA tile 64 x 64, K step 16, 4 warps (2 x 2), each warp a 32 x 32 block = 4 x 4 DMMA tiles (32 accumulator doubles
// per lane); operands are staged through shared memory with cp.async (two stages), laid out so that the global reads
// are contiguous whichever operand is transposed and the fragment reads hit the 2-wavefront minimum.
// Operand layouts (element (mn, k) of the 64 x 16 operand tile, mn = row of A or column of B):
//   LAY_MN   g[mn + ld*k]             contiguous along mn
//   LAY_K    g[k + ld*mn]             contiguous along k
//   LAY_SYM  g[pk(mn, k % nsym) + ld*(k / nsym)]   a stack of symmetric matrices stored packed (lower triangle, column-
//            major: pk(a>=b) = b*(2*nsym-b+1)/2 + a-b), slab stride ld: the RI tensor B[mu nu, Q] read as B^Q[mu, nu]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qlb {
namespace dmma {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int LDM = BM + 8;  // [k][m] layout: row stride 72 doubles (k rows land 16 banks apart)
constexpr int LDK = BK + 4;  // [m][k] layout: row stride 20 doubles
constexpr int STAGE_DOUBLES = (BK * LDM > BM * LDK) ? BK * LDM : BM * LDK;
constexpr int LAY_MN = 0, LAY_K = 1, LAY_SYM = 2;

__host__ __device__ __forceinline__ int64_t sym_pk(int64_t a, int64_t b, int64_t n) {  // packed index of (a,b), any order
  const int64_t hi = a > b ? a : b, lo = a > b ? b : a;
  return lo * (2 * n - lo + 1) / 2 + (hi - lo);
}

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

// one operand tile (64 x 16 of op(X)) into shared memory: [k][mn] for LAY_MN / LAY_SYM, [mn][k] for LAY_K
template <int LAY>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ g, int64_t ld, int nsym, int mn0, int k0, int MN, int K, int tid) {
  if (LAY == LAY_K) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + 128 * i, k = e % BK, mn = e / BK;
      const bool ok = (mn0 + mn < MN) && (k0 + k < K);
      cp_async8(s + mn * LDK + k, g + (ok ? (int64_t)(k0 + k) + ld * (mn0 + mn) : 0), ok);
    }
  } else {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + 128 * i, mn = e % BM, k = e / BM;
      const bool ok = (mn0 + mn < MN) && (k0 + k < K);
      int64_t off = 0;
      if (ok) {
        if (LAY == LAY_MN) off = (int64_t)(mn0 + mn) + ld * (k0 + k);
        else off = sym_pk(mn0 + mn, (k0 + k) % nsym, nsym) + ld * ((k0 + k) / nsym);
      }
      cp_async8(s + k * LDM + mn, g + off, ok);
    }
  }
}
template <int LAY>
__device__ __forceinline__ double frag(const double *s, int mn, int k) {
  return LAY == LAY_K ? s[mn * LDK + k] : s[k * LDM + mn];
}

// gridDim.z batches: operand pointers advance by strideA/strideB/strideC doubles per batch.
// LOWER_ONLY: skip CTA tiles strictly above the diagonal (symmetric results; the caller mirrors).
template <int LAYA, int LAYB, bool LOWER_ONLY>
__global__ void __launch_bounds__(128) dgemm_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int64_t lda,
                                                    int64_t strideA, const double *__restrict__ B, int64_t ldb, int64_t strideB,
                                                    double beta, double *C, int64_t ldc, int64_t strideC, int nsym) {
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
  load_tile<LAYA>(As[0], A, lda, nsym, m0, 0, M, K, tid);
  load_tile<LAYB>(Bs[0], B, ldb, nsym, n0, 0, N, K, tid);
  cp_async_commit();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<LAYA>(As[cur ^ 1], A, lda, nsym, m0, (kt + 1) * BK, M, K, tid);
      load_tile<LAYB>(Bs[cur ^ 1], B, ldb, nsym, n0, (kt + 1) * BK, N, K, tid);
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
      for (int i = 0; i < 4; ++i) a[i] = frag<LAYA>(As[cur], wm + 8 * i + fr, kk + fc);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = frag<LAYB>(Bs[cur], wn + 8 * j + fr, kk + fc);
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

struct Operand {
  const double *p;
  int64_t ld;      // leading dimension (LAY_MN / LAY_K) or slab stride (LAY_SYM)
  int lay;
  int64_t stride;  // per-batch advance
};

// general launcher.  Supported layout pairs: every combination of LAY_MN / LAY_K, plus (LAY_SYM, LAY_K|LAY_MN) and (LAY_MN, LAY_SYM).
inline cudaError_t dgemm_ex(cudaStream_t s, int M, int N, int K, double alpha, Operand A, Operand B, double beta, double *C, int64_t ldc,
                            int64_t strideC = 0, int batch = 1, bool lower_only = false, int nsym = 1) {
  if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
  const dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, batch), block(128);
  bool done = false;
#define QLB_DGEMM_CASE(LA, LB)                                                                                                        \
  if (!done && A.lay == LA && B.lay == LB) {                                                                                          \
    done = true;                                                                                                                      \
    if (lower_only)                                                                                                                   \
      dgemm_kernel<LA, LB, true><<<grid, block, 0, s>>>(M, N, K, alpha, A.p, A.ld, A.stride, B.p, B.ld, B.stride, beta, C, ldc, strideC, nsym); \
    else                                                                                                                              \
      dgemm_kernel<LA, LB, false><<<grid, block, 0, s>>>(M, N, K, alpha, A.p, A.ld, A.stride, B.p, B.ld, B.stride, beta, C, ldc, strideC, nsym); \
  }
  QLB_DGEMM_CASE(LAY_MN, LAY_K)
  QLB_DGEMM_CASE(LAY_MN, LAY_MN)
  QLB_DGEMM_CASE(LAY_K, LAY_K)
  QLB_DGEMM_CASE(LAY_K, LAY_MN)
  QLB_DGEMM_CASE(LAY_SYM, LAY_K)
  QLB_DGEMM_CASE(LAY_SYM, LAY_MN)
  QLB_DGEMM_CASE(LAY_MN, LAY_SYM)
#undef QLB_DGEMM_CASE
  if (!done) return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// BLAS-style wrapper: ta: op(A) = A^T (A stored K x M, lda >= K); tb: op(B) = B^T (B stored N x K, ldb >= N)
inline cudaError_t dgemm(cudaStream_t s, bool ta, bool tb, int M, int N, int K, double alpha, const double *A, int64_t lda,
                         const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int batch = 1, int64_t strideA = 0,
                         int64_t strideB = 0, int64_t strideC = 0, bool lower_only = false) {
  // A (M x K) is contiguous along m unless transposed; B (K x N) is contiguous along k unless transposed
  const Operand oa{A, lda, ta ? LAY_K : LAY_MN, strideA}, ob{B, ldb, tb ? LAY_MN : LAY_K, strideB};
  return dgemm_ex(s, M, N, K, alpha, oa, ob, beta, C, ldc, strideC, batch, lower_only);
}

}  // namespace dmma
}  // namespace qlb
