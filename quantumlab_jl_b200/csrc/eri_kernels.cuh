// eri_kernels.cuh -- the angular-momentum-class-specialised ERI + J/K digestion kernels.
//
// One template instantiation per canonical class (la>=lb, lc>=ld, bra class >= ket class).  A 2-D tile of
// (BY bra shell pairs) x (32 ket shell pairs) per CTA, one thread per shell quartet, one warp per bra pair:
// the bra side (shell indices, J_ab block, P_ab block) is warp-uniform, so J_ab is reduced with warp shuffles and
// written with one atomic per element per warp; the ket-indexed updates (J_cd, K_ac, K_ad, K_bc, K_bd) go to
// L2 with fire-and-forget FP64 atomics (RED.E.ADD.F64).
//
// Replaces contractTensorBlocksWithMatrix_CoulombLike/_ExchangeLike (SpecialMatricesModule.jl:33-100), which loop
// over all nsh^4 ordered quartets twice (J pass, K pass): here every unique quartet is visited once, screened by
// the integer Schwarz x density rule, and digested into J and K at the same time.
#pragma once
#include "md_core.cuh"

namespace qlb {

constexpr int TILE_KET = 32;  // = warp size: lanes of a warp share the bra pair
constexpr int TILE_BRA = 4;   // warps per CTA

// exact integer "quantised log": qlog(x) = ceil(8*log2(x)) evaluated without transcendental functions, so the
// host, the device and the CPU oracle agree bit for bit.  x = m*2^e, m in [1,2):  8e + #{j in 0..7 : m > 2^(j/8)}.
__host__ __device__ __forceinline__ int qlog_exact(double x) {
  if (!(x >= 1e-300)) return -32768;
  const double th[8] = {1.0,
                        1.0905077326652577,  // 2^(1/8)
                        1.189207115002721,   // 2^(2/8)
                        1.2968395546510096,  // 2^(3/8)
                        1.4142135623730951,  // 2^(4/8)
                        1.5422108254079407,  // 2^(5/8)
                        1.681792830507429,   // 2^(6/8)
                        1.8340080864093424}; // 2^(7/8)
  unsigned long long bits;
#ifdef __CUDA_ARCH__
  bits = (unsigned long long)__double_as_longlong(x);
#else
  __builtin_memcpy(&bits, &x, 8);
#endif
  const int e = (int)((bits >> 52) & 0x7ff) - 1023;
  unsigned long long mb = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
  double m;
#ifdef __CUDA_ARCH__
  m = __longlong_as_double((long long)mb);
#else
  __builtin_memcpy(&m, &mb, 8);
#endif
  int c = 0;
  for (int j = 0; j < 8; ++j) c += (m > th[j]) ? 1 : 0;
  return 8 * e + c;
}

struct ClassParams {
  // shells
  const double *__restrict__ sh_xyz;   // 3*nsh
  const int *__restrict__ sh_boff;     // first basis function of each shell
  int nsh, nbf;
  // shell pairs (sorted by class, then Schwarz bound descending)
  const int *__restrict__ pair_a;
  const int *__restrict__ pair_b;
  const int *__restrict__ pair_poff;   // npairs+1 primitive-pair offsets
  const int *__restrict__ pair_ql;     // qlog(Q_ab)
  PrimPairs pp;
  const double *__restrict__ boys;
  // this launch: class ranges
  int bra_off, bra_n, ket_off, ket_n, same;
  // screening
  const int *__restrict__ dl;          // nsh*nsh, qlog(max |P| over the shell block)
  int tlog, dlmax, screen;
  // sharding of bra tile rows over ranks, grid decoding
  int rank, nranks, nbx;
  // data
  const double *__restrict__ P;
  double *J, *K;
  double *tensor;                      // MODE 1: N^4 output
  unsigned long long *counters;        // [0] quartets kept, [1] primitive quartets
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct QuartetCtx {
  bool active;
  double deg;
  int fa, fb, fc, fd;  // first basis function of each shell
};

// ---- J/K digestion of acc[a,b,c,d-DLO], d in [DLO,DHI)
template <int LA, int LB, int LC, int LD, bool UNR, int DLO, int DHI>
__device__ __forceinline__ void digest_jk(const double *acc, const QuartetCtx &q, const ClassParams &prm) {
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC);
  const int N = prm.nbf;
  const double *__restrict__ P = prm.P;
  double Jab[NA * NB];
#pragma unroll(UNR ? 64 : 1)
  for (int i = 0; i < NA * NB; ++i) Jab[i] = 0.0;
  if (q.active) {
    int id = 0;
#pragma unroll(UNR ? 16 : 1)
    for (int dx = LD; dx >= 0; --dx) {
#pragma unroll(UNR ? 16 : 1)
      for (int dy = LD - dx; dy >= 0; --dy, ++id) {
        if (id < DLO || id >= DHI) continue;
        const double nd = nonaxial(LD, dx, dy, LD - dx - dy) * q.deg;
        const int gd = q.fd + id;
        double Kad[NA], Kbd[NB];
#pragma unroll(UNR ? 16 : 1)
        for (int i = 0; i < NA; ++i) Kad[i] = 0.0;
#pragma unroll(UNR ? 16 : 1)
        for (int i = 0; i < NB; ++i) Kbd[i] = 0.0;
        int ic = 0;
#pragma unroll(UNR ? 16 : 1)
        for (int cx = LC; cx >= 0; --cx) {
#pragma unroll(UNR ? 16 : 1)
          for (int cy = LC - cx; cy >= 0; --cy, ++ic) {
            const double ncd = nonaxial(LC, cx, cy, LC - cx - cy) * nd;
            const int gc = q.fc + ic;
            const double pcd = __ldg(P + gc + (size_t)N * gd);
            double jcd = 0.0;
            double Kac[NA], Kbc[NB];
#pragma unroll(UNR ? 16 : 1)
            for (int i = 0; i < NA; ++i) Kac[i] = 0.0;
#pragma unroll(UNR ? 16 : 1)
            for (int i = 0; i < NB; ++i) Kbc[i] = 0.0;
            int ib = 0;
#pragma unroll(UNR ? 16 : 1)
            for (int bx = LB; bx >= 0; --bx) {
#pragma unroll(UNR ? 16 : 1)
              for (int by = LB - bx; by >= 0; --by, ++ib) {
                const double nbcd = nonaxial(LB, bx, by, LB - bx - by) * ncd;
                const int gb = q.fb + ib;
                const double pbd = __ldg(P + gb + (size_t)N * gd);
                const double pbc = __ldg(P + gb + (size_t)N * gc);
                int ia = 0;
#pragma unroll(UNR ? 16 : 1)
                for (int ax = LA; ax >= 0; --ax) {
#pragma unroll(UNR ? 16 : 1)
                  for (int ay = LA - ax; ay >= 0; --ay, ++ia) {
                    const int ga = q.fa + ia;
                    const double I = acc[ia + NA * (ib + NB * (ic + NC * (id - DLO)))] *
                                     (nonaxial(LA, ax, ay, LA - ax - ay) * nbcd);
                    Jab[ia + NA * ib] = fma(I, pcd, Jab[ia + NA * ib]);
                    jcd = fma(I, __ldg(P + ga + (size_t)N * gb), jcd);
                    Kac[ia] = fma(I, pbd, Kac[ia]);
                    Kad[ia] = fma(I, pbc, Kad[ia]);
                    Kbc[ib] = fma(I, __ldg(P + ga + (size_t)N * gd), Kbc[ib]);
                    Kbd[ib] = fma(I, __ldg(P + ga + (size_t)N * gc), Kbd[ib]);
                  }
                }
              }
            }
            atomicAdd(prm.J + gc + (size_t)N * gd, 0.5 * jcd);
#pragma unroll(UNR ? 16 : 1)
            for (int i = 0; i < NA; ++i) atomicAdd(prm.K + (q.fa + i) + (size_t)N * gc, 0.25 * Kac[i]);
#pragma unroll(UNR ? 16 : 1)
            for (int i = 0; i < NB; ++i) atomicAdd(prm.K + (q.fb + i) + (size_t)N * gc, 0.25 * Kbc[i]);
          }
        }
#pragma unroll(UNR ? 16 : 1)
        for (int i = 0; i < NA; ++i) atomicAdd(prm.K + (q.fa + i) + (size_t)N * gd, 0.25 * Kad[i]);
#pragma unroll(UNR ? 16 : 1)
        for (int i = 0; i < NB; ++i) atomicAdd(prm.K + (q.fb + i) + (size_t)N * gd, 0.25 * Kbd[i]);
      }
    }
  }
  // J_ab: the bra pair is warp-uniform -> shuffle-reduce, one atomic per element per warp
  const int lane = threadIdx.x;
#pragma unroll(UNR ? 64 : 1)
  for (int i = 0; i < NA * NB; ++i) {
    const double s = warp_sum(Jab[i]);
    if (lane == 0 && s != 0.0) atomicAdd(prm.J + (q.fa + i % NA) + (size_t)N * (q.fb + i / NA), 0.5 * s);
  }
}

// ---- MODE 1: scatter the block into the full N^4 tensor at its (up to) 8 symmetric positions
template <int LA, int LB, int LC, int LD, int DLO, int DHI>
__device__ __forceinline__ void scatter_tensor(const double *acc, const QuartetCtx &q, const ClassParams &prm) {
  constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC);
  if (!q.active) return;
  const size_t N = prm.nbf;
  double *T = prm.tensor;
  int id = 0;
  for (int dx = LD; dx >= 0; --dx)
    for (int dy = LD - dx; dy >= 0; --dy, ++id) {
      if (id < DLO || id >= DHI) continue;
      const double nd = nonaxial(LD, dx, dy, LD - dx - dy);
      int ic = 0;
      for (int cx = LC; cx >= 0; --cx)
        for (int cy = LC - cx; cy >= 0; --cy, ++ic) {
          const double ncd = nonaxial(LC, cx, cy, LC - cx - cy) * nd;
          int ib = 0;
          for (int bx = LB; bx >= 0; --bx)
            for (int by = LB - bx; by >= 0; --by, ++ib) {
              const double nbcd = nonaxial(LB, bx, by, LB - bx - by) * ncd;
              int ia = 0;
              for (int ax = LA; ax >= 0; --ax)
                for (int ay = LA - ax; ay >= 0; --ay, ++ia) {
                  const double I = acc[ia + NA * (ib + NB * (ic + NC * (id - DLO)))] *
                                   (nonaxial(LA, ax, ay, LA - ax - ay) * nbcd);
                  const size_t a = q.fa + ia, b = q.fb + ib, c = q.fc + ic, d = q.fd + id;
                  T[a + N * (b + N * (c + N * d))] = I;
                  T[b + N * (a + N * (c + N * d))] = I;
                  T[a + N * (b + N * (d + N * c))] = I;
                  T[b + N * (a + N * (d + N * c))] = I;
                  T[c + N * (d + N * (a + N * b))] = I;
                  T[d + N * (c + N * (a + N * b))] = I;
                  T[c + N * (d + N * (b + N * a))] = I;
                  T[d + N * (c + N * (b + N * a))] = I;
                }
            }
        }
    }
}

template <int LA, int LB, int LC, int LD, bool UNR, int MODE, int DLO, int DSTEP>
struct PassRunner {
  static constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  static __device__ __forceinline__ void run(const PairRef &bra, const PairRef &ket, const QuartetCtx &q,
                                             const ClassParams &prm) {
    double acc[NA * NB * NC * DSTEP];
#pragma unroll(UNR ? 256 : 1)
    for (int i = 0; i < NA * NB * NC * DSTEP; ++i) acc[i] = 0.0;
    quartet_accumulate<LA, LB, LC, LD, UNR, DLO, DLO + DSTEP>(prm.pp, bra, ket, prm.boys, acc);
    if (MODE == 0) digest_jk<LA, LB, LC, LD, UNR, DLO, DLO + DSTEP>(acc, q, prm);
    else scatter_tensor<LA, LB, LC, LD, DLO, DLO + DSTEP>(acc, q, prm);
    if constexpr (DLO + DSTEP < ND) PassRunner<LA, LB, LC, LD, UNR, MODE, DLO + DSTEP, DSTEP>::run(bra, ket, q, prm);
  }
};

// MODE 0: J/K digestion with screening; MODE 1: tensor scatter (no screening).
// DSTEP: ket-d components per pass (ncart(LD) = single pass; 1 = one pass per d component, smaller accumulators).
template <int LA, int LB, int LC, int LD, bool UNR, int MODE, int DSTEP>
__global__ void __launch_bounds__(TILE_KET *TILE_BRA) eri_class_kernel(const __grid_constant__ ClassParams prm) {
  const int row = (blockIdx.x / prm.nbx) * prm.nranks + prm.rank;
  const int bx = blockIdx.x % prm.nbx;
  const int ib0 = row * TILE_BRA, ik0 = bx * TILE_KET;
  if (ib0 >= prm.bra_n) return;
  if (prm.same && ik0 > ib0 + TILE_BRA - 1) return;
  if (prm.screen && prm.pair_ql[prm.bra_off + ib0] + prm.pair_ql[prm.ket_off + ik0] + prm.dlmax < prm.tlog) return;

  const int ib = ib0 + threadIdx.y, ik = ik0 + threadIdx.x;
  const bool valid = ib < prm.bra_n && ik < prm.ket_n && (!prm.same || ik <= ib);
  const int gb = prm.bra_off + min(ib, prm.bra_n - 1), gk = prm.ket_off + min(ik, prm.ket_n - 1);
  const int sa = prm.pair_a[gb], sb = prm.pair_b[gb], sc = prm.pair_a[gk], sd = prm.pair_b[gk];
  bool active = valid;
  if (prm.screen && valid) {
    const int s = prm.pair_ql[gb] + prm.pair_ql[gk];
    if (s + prm.dlmax < prm.tlog) active = false;
    else {
      const int *__restrict__ dl = prm.dl;
      const int n = prm.nsh;
      int dm = max(dl[sa + n * sb], dl[sc + n * sd]);
      dm = max(dm, max(dl[sa + n * sc], dl[sa + n * sd]));
      dm = max(dm, max(dl[sb + n * sc], dl[sb + n * sd]));
      active = (s + dm >= prm.tlog);
    }
  }
  PairRef bra, ket;
  bra.off = prm.pair_poff[gb]; bra.n = active ? prm.pair_poff[gb + 1] - bra.off : 0;
  ket.off = prm.pair_poff[gk]; ket.n = active ? prm.pair_poff[gk + 1] - ket.off : 0;
  {
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (m == 0u) return;  // warp-uniform
    const unsigned np = __reduce_add_sync(0xffffffffu, (unsigned)(bra.n * ket.n));
    if (threadIdx.x == 0 && prm.counters) {
      atomicAdd(prm.counters + 0, (unsigned long long)__popc(m));
      atomicAdd(prm.counters + 1, (unsigned long long)np);
    }
  }
  bra.Ax = prm.sh_xyz[3 * sa]; bra.Ay = prm.sh_xyz[3 * sa + 1]; bra.Az = prm.sh_xyz[3 * sa + 2];
  bra.Bx = prm.sh_xyz[3 * sb]; bra.By = prm.sh_xyz[3 * sb + 1]; bra.Bz = prm.sh_xyz[3 * sb + 2];
  ket.Ax = prm.sh_xyz[3 * sc]; ket.Ay = prm.sh_xyz[3 * sc + 1]; ket.Az = prm.sh_xyz[3 * sc + 2];
  ket.Bx = prm.sh_xyz[3 * sd]; ket.By = prm.sh_xyz[3 * sd + 1]; ket.Bz = prm.sh_xyz[3 * sd + 2];
  QuartetCtx q;
  q.active = active;
  q.deg = ((sa == sb) ? 1.0 : 2.0) * ((sc == sd) ? 1.0 : 2.0) * ((prm.same && ib == ik) ? 1.0 : 2.0);
  if (MODE == 1) q.deg = 1.0;
  q.fa = prm.sh_boff[sa]; q.fb = prm.sh_boff[sb]; q.fc = prm.sh_boff[sc]; q.fd = prm.sh_boff[sd];
  PassRunner<LA, LB, LC, LD, UNR, MODE, 0, DSTEP>::run(bra, ket, q, prm);
}

// ---- Schwarz bounds: one thread per shell pair, Q = sqrt(max_ab |(ab|ab)|)
struct SchwarzParams {
  const double *__restrict__ sh_xyz;
  const int *__restrict__ pair_a;
  const int *__restrict__ pair_b;
  const int *__restrict__ pair_poff;
  PrimPairs pp;
  const double *__restrict__ boys;
  int off, n;
  double *Q;  // per pair
};

template <int LA, int LB>
__global__ void __launch_bounds__(64) schwarz_kernel(const __grid_constant__ SchwarzParams prm) {
  constexpr int NA = ncart(LA), NB = ncart(LB);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= prm.n) return;
  const int g = prm.off + i;
  const int sa = prm.pair_a[g], sb = prm.pair_b[g];
  PairRef pr;
  pr.off = prm.pair_poff[g]; pr.n = prm.pair_poff[g + 1] - pr.off;
  pr.Ax = prm.sh_xyz[3 * sa]; pr.Ay = prm.sh_xyz[3 * sa + 1]; pr.Az = prm.sh_xyz[3 * sa + 2];
  pr.Bx = prm.sh_xyz[3 * sb]; pr.By = prm.sh_xyz[3 * sb + 1]; pr.Bz = prm.sh_xyz[3 * sb + 2];
  double acc[NA * NB * NA * NB];
  for (int k = 0; k < NA * NB * NA * NB; ++k) acc[k] = 0.0;
  quartet_accumulate<LA, LB, LA, LB, false, 0, NB>(prm.pp, pr, pr, prm.boys, acc);
  double m = 0.0;
  int ib = 0;
  for (int bx = LB; bx >= 0; --bx)
    for (int by = LB - bx; by >= 0; --by, ++ib) {
      const double nb = nonaxial(LB, bx, by, LB - bx - by);
      int ia = 0;
      for (int ax = LA; ax >= 0; --ax)
        for (int ay = LA - ax; ay >= 0; --ay, ++ia) {
          const double na = nonaxial(LA, ax, ay, LA - ax - ay);
          const double v = fabs(acc[ia + NA * (ib + NB * (ia + NA * ib))]) * (na * nb) * (na * nb);
          m = fmax(m, v);
        }
    }
  prm.Q[g] = sqrt(m);
}

}  // namespace qlb
