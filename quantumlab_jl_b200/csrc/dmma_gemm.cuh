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
#include <stdlib.h>
#include <algorithm>

namespace qlb {
namespace dmma {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int GK = BK;  // K step of both tile sizes
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

template <int LAY, int ROWS, int NTHR, int SLDM, int SLDK>
__device__ __forceinline__ void load_tile_g(double *s, const double *__restrict__ g, int64_t ld, int nsym, int mn0, int k0, int MN, int K, int tid);
template <int LAY>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ g, int64_t ld, int nsym, int mn0, int k0, int MN, int K, int tid) {
  load_tile_g<LAY, BM, 128, LDM, LDK>(s, g, ld, nsym, mn0, k0, MN, K, tid);
}
template <int LAY>
__device__ __forceinline__ double frag(const double *s, int mn, int k) {
  return LAY == LAY_K ? s[mn * LDK + k] : s[k * LDM + mn];
}

// gridDim.z batches: operand pointers advance by strideA/strideB/strideC doubles per batch.
// LOWER_ONLY: skip CTA tiles strictly above the diagonal (symmetric results; the caller mirrors).
// kper > 0: split-K -- gridDim.z slices of kper (a multiple of the K step) columns of the contraction index instead of
// batches; every slice adds alpha * its partial product into C with RED.ADD.F64 (the launcher has applied beta already).
template <int LAYA, int LAYB, bool LOWER_ONLY>
__global__ void __launch_bounds__(128) dgemm_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int64_t lda,
                                                    int64_t strideA, const double *__restrict__ B, int64_t ldb, int64_t strideB,
                                                    double beta, double *C, int64_t ldc, int64_t strideC, int nsym, int kper) {
  __shared__ double As[2][STAGE_DOUBLES], Bs[2][STAGE_DOUBLES];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (LOWER_ONLY && n0 > m0 + BM - 1) return;
  int kbeg = 0;
  if (kper > 0) {
    kbeg = blockIdx.z * kper;
    K = min(K, kbeg + kper);
  } else {
    A += strideA * blockIdx.z;
    B += strideB * blockIdx.z;
    C += strideC * blockIdx.z;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int fr = lane >> 2, fc = lane & 3;  // fragment row / column
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nk = (K - kbeg + BK - 1) / BK;
  load_tile<LAYA>(As[0], A, lda, nsym, m0, kbeg, M, K, tid);
  load_tile<LAYB>(Bs[0], B, ldb, nsym, n0, kbeg, N, K, tid);
  cp_async_commit();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile<LAYA>(As[cur ^ 1], A, lda, nsym, m0, kbeg + (kt + 1) * BK, M, K, tid);
      load_tile<LAYB>(Bs[cur ^ 1], B, ldb, nsym, n0, kbeg + (kt + 1) * BK, N, K, tid);
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
          if (kper > 0) atomicAdd(c, alpha * acc[i][j][h]);
          else *c = (beta == 0.0) ? alpha * acc[i][j][h] : alpha * acc[i][j][h] + beta * *c;
        }
      }
}

// C *= beta (beta == 0: C = 0) ahead of a split-K product
static __global__ void scale_kernel(int M, int N, double beta, double *C, int64_t ldc) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (m < M && n < N) {
    double *c = C + (int64_t)m + ldc * n;
    *c = (beta == 0.0) ? 0.0 : beta * *c;
  }
}

// ------------------------------------------------------------------ large tiles: 128 x 128 x 16, 8 warps, 3 cp.async stages
// The 64 x 64 kernel above moves (64+64)*8 B of operands per 64*64*2 flops = 8 flop/B through L2 -> shared memory: at the
// measured FP64 rate that is ~4.6 TB/s of L2 traffic, and it ran at 0.49 of the DMMA peak on the RI-K contractions
// (profiles/r02).  This kernel doubles the intensity (16 flop/B): CTA tile 128 x 128, 8 warps as 2 (M) x 4 (N), warp tile
// 64 x 32 = 8 x 4 DMMA tiles (64 accumulator doubles per lane), three cp.async stages (one barrier per K step), the
// packed-symmetric operand's slab / column decomposition hoisted out of the element loop.
constexpr int GM = 128, GSTAGES = 3;  // GN = 32 * NT (template): 128 wide, or 64 wide for awkward column counts
constexpr int GLDM = GM + 8;  // [k][mn] layout
constexpr int GLDK = GK + 4;  // [mn][k] layout
constexpr int GSTAGE_DOUBLES = (GK * GLDM > GM * GLDK) ? GK * GLDM : GM * GLDK;
constexpr size_t GSMEM_BYTES = (size_t)2 * GSTAGES * GSTAGE_DOUBLES * sizeof(double);  // B stages sized like A's (GN <= GM)

// one operand tile (ROWS x 16 of op(X)) into shared memory with NTHR threads; [k][mn] (row stride SLDM) for LAY_MN / LAY_SYM,
// [mn][k] (row stride SLDK) for LAY_K.  The packed-symmetric operand's slab / column decomposition (one integer division) is
// done once per tile and advanced incrementally (ncu: with a division per element the index arithmetic of X = B^Q C
// out-issued the DMMAs).
template <int LAY, int ROWS, int NTHR, int SLDM, int SLDK>
__device__ __forceinline__ void load_tile_g(double *s, const double *__restrict__ g, int64_t ld, int nsym, int mn0, int k0, int MN, int K, int tid) {
  if (LAY == LAY_K) {
#pragma unroll
    for (int i = 0; i < (ROWS * GK) / NTHR; ++i) {
      const int e = tid + NTHR * i, k = e % GK, mn = e / GK;
      const bool ok = (mn0 + mn < MN) && (k0 + k < K);
      cp_async8(s + mn * SLDK + k, g + (ok ? (int64_t)(k0 + k) + ld * (mn0 + mn) : 0), ok);
    }
  } else {
    constexpr int KSTEP = NTHR / ROWS;
    const int mn = tid % ROWS, kb = tid / ROWS;  // this thread: row mn, k = kb, kb+KSTEP, ...
    const bool mok = mn0 + mn < MN;
    const int64_t a = mn0 + mn;
    int slab = 0, kk = k0 + kb;
    if (LAY == LAY_SYM) {
      slab = kk / nsym;
      kk -= slab * nsym;
    }
#pragma unroll
    for (int i = 0; i < (ROWS * GK) / NTHR; ++i) {
      const int k = kb + KSTEP * i;
      const bool ok = mok && (k0 + k < K);
      int64_t off = 0;
      if (LAY == LAY_MN) {
        if (ok) off = a + ld * (k0 + k);
      } else {
        if (ok) {
          const int64_t b = kk, hi = a > b ? a : b, lo = a > b ? b : a;
          off = lo * (2 * (int64_t)nsym - lo + 1) / 2 + (hi - lo) + ld * slab;
        }
        kk += KSTEP;
        while (kk >= nsym) {
          kk -= nsym;
          ++slab;
        }
      }
      cp_async8(s + k * SLDM + mn, g + off, ok);
    }
  }
}
template <int LAY, int ROWS>
__device__ __forceinline__ void load_tile_big(double *s, const double *__restrict__ g, int64_t ld, int nsym, int mn0, int k0, int MN, int K, int tid) {
  load_tile_g<LAY, ROWS, 256, GLDM, GLDK>(s, g, ld, nsym, mn0, k0, MN, K, tid);
}
template <int LAY>
__device__ __forceinline__ double frag_big(const double *s, int mn, int k) {
  return LAY == LAY_K ? s[mn * GLDK + k] : s[k * GLDM + mn];
}

template <int LAYA, int LAYB, bool LOWER_ONLY, int NT>
__global__ void __launch_bounds__(256, 1) dgemm_big_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int64_t lda,
                                                           int64_t strideA, const double *__restrict__ B, int64_t ldb, int64_t strideB,
                                                           double beta, double *C, int64_t ldc, int64_t strideC, int nsym, int kper) {
  extern __shared__ __align__(16) double g_smem[];
  double *As = g_smem, *Bs = g_smem + GSTAGES * GSTAGE_DOUBLES;
  constexpr int GN = 32 * NT;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  if (LOWER_ONLY && n0 > m0 + GM - 1) return;
  int kbeg = 0;
  if (kper > 0) {
    kbeg = blockIdx.z * kper;
    K = min(K, kbeg + kper);
  } else {
    A += strideA * blockIdx.z;
    B += strideB * blockIdx.z;
    C += strideC * blockIdx.z;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * (8 * NT);
  const int fr = lane >> 2, fc = lane & 3;
  double acc[8][NT][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nk = (K - kbeg + GK - 1) / GK;
#pragma unroll
  for (int st = 0; st < GSTAGES - 1; ++st) {
    if (st < nk) {
      load_tile_big<LAYA, GM>(As + st * GSTAGE_DOUBLES, A, lda, nsym, m0, kbeg + st * GK, M, K, tid);
      load_tile_big<LAYB, GN>(Bs + st * GSTAGE_DOUBLES, B, ldb, nsym, n0, kbeg + st * GK, N, K, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();  // stage kt has landed (this thread's copies) ...
    __syncthreads();               // ... and everyone's; everyone is also done reading stage kt-1
    {
      const int nx = kt + GSTAGES - 1;  // refill the stage consumed at step kt-1
      if (nx < nk) {
        load_tile_big<LAYA, GM>(As + (nx % GSTAGES) * GSTAGE_DOUBLES, A, lda, nsym, m0, kbeg + nx * GK, M, K, tid);
        load_tile_big<LAYB, GN>(Bs + (nx % GSTAGES) * GSTAGE_DOUBLES, B, ldb, nsym, n0, kbeg + nx * GK, N, K, tid);
      }
      cp_async_commit();
    }
    const double *as = As + (kt % GSTAGES) * GSTAGE_DOUBLES, *bs = Bs + (kt % GSTAGES) * GSTAGE_DOUBLES;
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double a[8], b[NT];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = frag_big<LAYA>(as, wm + 8 * i + fr, kk + fc);
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = frag_big<LAYB>(bs, wn + 8 * j + fr, kk + fc);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = m0 + wm + 8 * i + fr, n = n0 + wn + 8 * j + 2 * fc + h;
        if (m < M && n < N) {
          double *c = C + (int64_t)m + ldc * n;
          if (kper > 0) atomicAdd(c, alpha * acc[i][j][h]);
          else *c = (beta == 0.0) ? alpha * acc[i][j][h] : alpha * acc[i][j][h] + beta * *c;
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
  bool done = false;
  // split-K: an unbatched product whose tile grid cannot fill the 148 SMs but whose contraction index is long (K += X X^T
  // of the RI exchange: 15 tiles of 128 x 128 at N = 608 against K = nocc * Q ~ 1e5) is cut along K into gridDim.z slices
  auto plan_split = [&](int tiles, int kstep) -> int {
    if (batch != 1 || tiles >= 2 * 148) return 0;
    const int nk = (K + kstep - 1) / kstep;
    int splits = std::min((2 * 148 + tiles - 1) / tiles, nk / 8);
    if (splits < 2) return 0;
    const int per = ((nk + splits - 1) / splits) * kstep;  // a multiple of the K step
    return per;
  };
  auto pre_scale = [&]() -> cudaError_t {
    scale_kernel<<<dim3((M + 127) / 128, N), 128, 0, s>>>(M, N, beta, C, ldc);
    return cudaGetLastError();
  };
  // large tiles (128 x 128, or 128 x 64 when that pads the columns less) once the problem fills them; small problems
  // and strongly padded ones stay on the 64 x 64 kernel
  // measured (profiles/r02, gpurun r2o/r2p): the 128 x 128 tiles win on the plain-layout products (K += X X^T: tensor pipe busy
  // 73 % of the cycles against 62 %; RI build 36.2 vs 41.4 ms at N = 608), the 128 x 64 variant on the packed-symmetric
  // operand loses (X = B^Q C at N = 1600: 1108 vs 849 ms per build) -- so: wide tiles, no packed operand, little padding
  static const bool no_big = getenv("QLB_DGEMM_SMALL") != nullptr, force_big = getenv("QLB_DGEMM_BIG") != nullptr;
  const int64_t mpad = (int64_t)((M + GM - 1) / GM) * GM;
  const int64_t npad128 = (int64_t)((N + 127) / 128) * 128, npad64 = (int64_t)((N + 63) / 64) * 64;
  const bool wide = force_big ? 16 * npad128 <= 17 * npad64 : true;
  const bool plain = A.lay != LAY_SYM && B.lay != LAY_SYM;
  if (!no_big && (force_big || plain) && M >= GM && N >= (wide ? 128 : 64) && 4 * mpad * (wide ? npad128 : npad64) <= 5 * (int64_t)M * N) {
    const int gn = wide ? 128 : 64;
    dim3 ggrid((M + GM - 1) / GM, (N + gn - 1) / gn, batch);
    const int kper = plan_split(lower_only ? (int)(ggrid.x * (ggrid.y + 1) / 2) : (int)(ggrid.x * ggrid.y), GK);
    if (kper > 0) {
      ggrid.z = (K + kper - 1) / kper;
      if (cudaError_t pe = pre_scale()) return pe;
    }
#define QLB_DGEMM_BIG(LA, LB)                                                                                                         \
  if (!done && A.lay == LA && B.lay == LB) {                                                                                          \
    done = true;                                                                                                                      \
    auto kern = wide ? (lower_only ? dgemm_big_kernel<LA, LB, true, 4> : dgemm_big_kernel<LA, LB, false, 4>)                          \
                     : (lower_only ? dgemm_big_kernel<LA, LB, true, 2> : dgemm_big_kernel<LA, LB, false, 2>);                         \
    cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GSMEM_BYTES);                       \
    if (ae != cudaSuccess) return ae;                                                                                                 \
    kern<<<ggrid, 256, GSMEM_BYTES, s>>>(M, N, K, alpha, A.p, A.ld, A.stride, B.p, B.ld, B.stride, beta, C, ldc, strideC, nsym, kper); \
  }
    QLB_DGEMM_BIG(LAY_MN, LAY_K)
    QLB_DGEMM_BIG(LAY_MN, LAY_MN)
    QLB_DGEMM_BIG(LAY_K, LAY_K)
    QLB_DGEMM_BIG(LAY_K, LAY_MN)
    QLB_DGEMM_BIG(LAY_SYM, LAY_K)
    QLB_DGEMM_BIG(LAY_SYM, LAY_MN)
    QLB_DGEMM_BIG(LAY_MN, LAY_SYM)
#undef QLB_DGEMM_BIG
    if (!done) return cudaErrorInvalidValue;
    return cudaGetLastError();
  }
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, batch);
  const dim3 block(128);
  const int kper = plan_split(lower_only ? (int)(grid.x * (grid.y + 1) / 2) : (int)(grid.x * grid.y), BK);
  if (kper > 0) {
    grid.z = (K + kper - 1) / kper;
    if (cudaError_t pe = pre_scale()) return pe;
  }
#define QLB_DGEMM_CASE(LA, LB)                                                                                                        \
  if (!done && A.lay == LA && B.lay == LB) {                                                                                          \
    done = true;                                                                                                                      \
    if (lower_only)                                                                                                                   \
      dgemm_kernel<LA, LB, true><<<grid, block, 0, s>>>(M, N, K, alpha, A.p, A.ld, A.stride, B.p, B.ld, B.stride, beta, C, ldc, strideC, nsym, kper); \
    else                                                                                                                              \
      dgemm_kernel<LA, LB, false><<<grid, block, 0, s>>>(M, N, K, alpha, A.p, A.ld, A.stride, B.p, B.ld, B.stride, beta, C, ldc, strideC, nsym, kper); \
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
