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
constexpr int DL_SMEM_MAX_NSH = 2048;  // above this the screening reads dl from global memory

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
  // shell pairs (sorted by class, then coarse Schwarz-bound bin descending, then shell indices: see api.cu)
  const int *__restrict__ pair_a;
  const int *__restrict__ pair_b;
  const int *__restrict__ pair_poff;   // per pair: double2 index of field 0 of its first primitive record (tile layout)
  const int *__restrict__ pair_np;     // per pair: number of primitive pairs
  const int *__restrict__ pair_ql;     // qlog(Q_ab)
  const int *__restrict__ pair_qls;    // max of pair_ql over this and every LATER pair of the class (non-increasing: ends walks)
  const int *__restrict__ pair_dl;     // per build: qlog(max |P|) over the pair's own shell block (coalesced in the walk)
  // packed per-pair record, 5 x 16 B (field-major in blocks of 32 pairs, see load_pair_info): {Ax,Ay} {Az,Bx} {By,Bz} int4{a,b,poff,np} int4{fa,fb,ql,0}: what a batch needs about a
  // ket pair arrives with 5 vector loads instead of 14 scattered scalar ones (ncu: the light classes are bound by L1
  // wavefronts of exactly those gathers)
  const double2 *__restrict__ pair_info;
  int dl_smem;                         // 1: the kernels stage the bra shells' rows of dl in shared memory (nsh <= QLB_DL_SMEM_MAX_NSH)
  PrimPairs pp;
  const double *__restrict__ boys;
  // this launch: class ranges
  int bra_off, bra_n, ket_off, ket_n, same;
  // screening
  const int *__restrict__ dl;          // nsh*nsh, qlog(max |P| over the shell block)
  int tlog, screen;                    // the density bound max qlog|P| is read from dl[nsh*nsh] on the device
  // sharding of bra tile rows over ranks, grid decoding
  int rank, nranks, nbx;
  int ysplit;                          // CTAs per bra tile row; CTA y handles ket tiles y, y+ysplit, ...
  // data
  const double *__restrict__ P;
  double *J, *K;
  double *tensor;                      // MODE 1: N^4 output; MODE 2: B3 (N*N*Naux) or M (Naux*Naux)
  int ri_mode, ri_n, ri_naux, ri_aux_fn0;  // MODE 2 (see scatter_ri)
  unsigned long long *counters;        // [0] quartets kept, [1] primitive quartets
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct PairInfo {  // unpacked pair_info record
  double Ax, Ay, Az, Bx, By, Bz;
  int a, b, poff, np, fa, fb, ql;
};
// Layout: blocks of 32 consecutive pairs, field-major inside a block ([block][field 0..4][pair]), so that the survivors of
// a batch -- which come from two or three ket tiles -- share cache lines per field load (80-byte records per pair made
// every lane of the five loads touch its own line: ncu, 160 L1 wavefronts per batch, as many as the primitive loop of a
// K = 10 batch).
constexpr int PAIR_INFO_BLOCK = 5 * 32;  // double2 per block of 32 pairs
__device__ __forceinline__ PairInfo load_pair_info(const double2 *__restrict__ info, int g) {
  const double2 *r = info + (size_t)(g >> 5) * PAIR_INFO_BLOCK + (g & 31);
  const double2 r0 = __ldg(r), r1 = __ldg(r + 32), r2 = __ldg(r + 64);
  const int4 i3 = __ldg(reinterpret_cast<const int4 *>(r + 96)), i4 = __ldg(reinterpret_cast<const int4 *>(r + 128));
  PairInfo o;
  o.Ax = r0.x; o.Ay = r0.y; o.Az = r1.x; o.Bx = r1.y; o.By = r2.x; o.Bz = r2.y;
  o.a = i3.x; o.b = i3.y; o.poff = i3.z; o.np = i3.w; o.fa = i4.x; o.fb = i4.y; o.ql = i4.z;
  return o;
}

struct QuartetCtx {
  bool active;
  double deg;
  int fa, fb, fc, fd;  // first basis function of each shell
};

}  // namespace qlb

// The ERI / digestion templates are compiled twice from the same source: namespace `unr` with every loop fully
// unrolled (register-resident intermediates) and namespace `rol` with rolled loops over local-memory arrays.
// (A plain `#pragma unroll` is required for the unrolled flavour: `#pragma unroll N` on the triangular inner loops
// makes the compiler unroll-by-N with remainder loops before the trip counts are known.)
#define QLB_U _Pragma("unroll")
#define QLB_NS unr
#include "eri_body.inc"
#undef QLB_U
#undef QLB_NS
#define QLB_U _Pragma("unroll 1")
#define QLB_NS rol
#include "eri_body.inc"
#undef QLB_U
#undef QLB_NS

namespace qlb {

// ---- Schwarz bounds: one thread per shell pair, Q = sqrt(max_ab |(ab|ab)|)
struct SchwarzParams {
  const double *__restrict__ sh_xyz;
  const int *__restrict__ pair_a;
  const int *__restrict__ pair_b;
  const int *__restrict__ pair_poff;
  const int *__restrict__ pair_np;
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
  pr.off = prm.pair_poff[g]; pr.n = prm.pair_np[g];
  pr.Ax = prm.sh_xyz[3 * sa]; pr.Ay = prm.sh_xyz[3 * sa + 1]; pr.Az = prm.sh_xyz[3 * sa + 2];
  pr.Bx = prm.sh_xyz[3 * sb]; pr.By = prm.sh_xyz[3 * sb + 1]; pr.Bz = prm.sh_xyz[3 * sb + 2];
  double acc[NA * NB * NA * NB];
  for (int k = 0; k < NA * NB * NA * NB; ++k) acc[k] = 0.0;
  rol::quartet_accumulate<LA, LB, LA, LB, 0, NB>(prm.pp, pr, pr, prm.boys + (size_t)(2 * (LA + LB)) * BOYS_TABLE_DOUBLES, acc);
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
