// qlb_internal.h -- host-side structures shared by the translation units of libqlb200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/qlb200.h"
#include "eri_kernels.cuh"

namespace qlb {

constexpr int NCLASS = QLB_NCLASS;    // ss ps pp ds dp dd
constexpr int NQCLASS = QLB_NQCLASS;  // X>=Y pairs of the above
constexpr int NCLASS_AUX = 7;         // (auxiliary shell, unit s) pair classes of the RI path: ss ps . ds . . fs -> s, p, d, f shells
inline int class_of(int la, int lb) { return la * (la + 1) / 2 + lb; }
inline void class_L(int c, int &la, int &lb) {
  la = 0;
  while ((la + 1) * (la + 2) / 2 <= c) ++la;
  lb = c - la * (la + 1) / 2;
}
inline int qclass_of(int X, int Y) { return X * (X + 1) / 2 + Y; }

struct Engine {
  bool ready = false;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  // class kernels are independent (they only add into J/K): they are dealt over the build stream plus these side
  // streams so that the tail of one class overlaps the body of the next
  static constexpr int NSIDE = 39;  // streams created; nside of them are used (QLB_NSIDE)
  int nside = 7;
  cudaStream_t side_stream[NSIDE] = {nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {nullptr};
  cudaStream_t copy_stream = nullptr;  // host -> device copies that overlap a build (the core Hamiltonian of qlb_fock)
  cudaEvent_t ev_copy = nullptr;
  double *d_boys = nullptr;
  bool profiling = false;
  // tuning knobs, read from the environment once in qlb_init
  // CTAs per SM per class launch that size ysplit.  Measured (gpurun r2v, B200): 29.3 / 30.8 / 32.7 / 34.2 ms per (H2O)32 build and
  // 79.0 / 81.7 / 84.9 / 87.1 ms per C40H82 build for 32 / 64 / 128 / 256 -- shorter walks end in more partial batches
  // 0 = the per-class table ys_target_of() in api.cu (default); QLB_YS_TARGET sets one value for every class
  int ys_target = 0, ys_tiles = 8;
  int dlsmem_min_tiles = 12;
  int launch_order = -1;  // class launch order: 0 high L first, 1 low L first, 2 expected cost descending, -1 auto (QLB_LAUNCH_ORDER)  // ket tiles per walk from which the kernels stage the density-bound rows in shared memory
  // width (in qlog units, 8 per factor 2) of the Schwarz-bound bins of the pair order; 1 = exact order of the bounds.
  // Measured on (H2O)32: wider bins (index-local tiles) LOSE -- 45.5 / 54 / 62 / 74 / 85 ms per build for 1 / 8 / 16 / 32 / 64
  int bin_width = 1;
  bool ysplit_full = false, no_schedule = false, no_dl_smem = false;
  // NCCL (loaded lazily with dlopen)
  void *nccl_lib = nullptr;
  void *nccl_comm = nullptr;
  int rank = 0, nranks = 1;
};
Engine &engine();            // the engine (device) the calling thread currently works on
void use_engine(Engine *e);  // switch to a handle's engine (and its device)
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define QLB_CUDA(call)                                   \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return cuda_fail(_e, #call);  \
  } while (0)

}  // namespace qlb

struct qlb_basis {
  qlb::Engine *eng = nullptr;  // the engine (device) that owns the device memory below
  int nsh = 0, nbf = 0, nprim_total = 0;
  std::vector<double> h_xyz, h_exps, h_coefs;
  std::vector<int> h_L, h_nprim, h_poff, h_boff;
  // pairs, sorted by class then Schwarz bound (descending)
  int64_t npairs = 0;
  std::vector<int> h_pair_a, h_pair_b, h_pair_poff, h_pair_np, h_pair_ql, h_pair_qls;
  std::vector<double> h_pair_Q;
  std::vector<int> h_kept_off, h_kept;  // kept primitive pairs per shell pair (CSR; code = ia * nprim_b + ib), see api.cu
  int *d_kept_off = nullptr, *d_kept = nullptr;
  int class_off[qlb::NCLASS + 1] = {0};
  // resolution of the identity: shells [0,n_orb) orbital, [n_orb, nsh-1) auxiliary, shell nsh-1 the unit s function;
  // pair group 0 = orbital pairs (class_off), group 1 = (aux, unit) pairs (ri_class_off).  n_orb == 0: ordinary basis
  int n_orb = 0;
  int ri_class_off[qlb::NCLASS_AUX + 1] = {0};
  // segments: each class is split by contraction degree (primitive pairs per shell pair) so that the threads of a
  // warp run primitive loops of similar length; every segment is sorted by Schwarz bound, descending
  std::vector<int> seg_off, seg_cls;
  int qlmax = -32768;
  // device
  double *d_xyz = nullptr, *d_exps = nullptr, *d_coefs = nullptr;
  int *d_poff = nullptr, *d_boff = nullptr, *d_L = nullptr;
  int *d_pair_a = nullptr, *d_pair_b = nullptr, *d_pair_poff = nullptr, *d_pair_np = nullptr, *d_pair_ql = nullptr, *d_pair_qls = nullptr;
  double *d_pair_Q = nullptr;
  double2 *d_pp = nullptr;  // 3 double2 per primitive pair: {p,Px} {Py,Pz} {k/p, 1/(2p)}
  double2 *d_pair_info = nullptr;  // 5 double2 per pair: packed centres / indices (eri_kernels.cuh: load_pair_info)
  int *d_pair_dl = nullptr;        // per build: density bound of the pair's own shell block
  int64_t nprimpairs = 0;
  // per-build scratch
  double *d_P = nullptr, *d_JK = nullptr, *d_H = nullptr;  // N*N, 2*N*N, N*N
  int *d_dl = nullptr;                                      // nsh*nsh (+1 slot for the max)
  unsigned long long *d_counters = nullptr;                 // 2 per quartet class
  cudaEvent_t ev[2 * qlb::NQCLASS + 4] = {nullptr};
  // multi-rank schedule feedback: all-rank per-class counters of the previous build (summed with a second, tiny
  // all-reduce next to the [J;K] one) drive a cost model that deals whole pieces of classes to ranks
  // [0, 2*NQCLASS): quartets / primitive quartets per class; [2*NQCLASS, 3*NQCLASS): class kernel times (ms) of the
  // calibration build.  They travel in the tail of the [J;K] all-reduce buffer (QLB_JK_TAIL doubles): one collective.
  double *h_gcounts = nullptr;       // pinned host copy of the all-rank sums
  double *d_ms_local = nullptr;      // staging of this rank's class times
  cudaEvent_t ev_gcounts = nullptr;
  bool gcounts_pending = false;
  int gcounts_nranks = 0;
  double calib_ms[qlb::NQCLASS] = {0};    // all-rank class times and work of the calibration build
  double calib_work[qlb::NQCLASS] = {0};
  int calib_nranks = 0;                   // 0: not calibrated
  double h_ms_local[qlb::NQCLASS] = {0};  // this rank's class times of the calibration build, until they are reduced
  bool ms_local_pending = false;
  // density bound of the previous build (pinned; sizes the grids of the next build, never decides a quartet)
  int *h_dlmax = nullptr;
  cudaEvent_t ev_dlmax = nullptr;
  bool dlmax_pending = false;
};

namespace qlb {
// class kernel launchers (class_kernels.cu)
int launch_class_kernel(int X, int Y, int mode, unsigned grid, cudaStream_t s, const ClassParams &prm);
int launch_schwarz_kernel(int X, unsigned grid, cudaStream_t s, const SchwarzParams &prm);
// flops per primitive quartet / digestion flops per contracted quartet for a quartet class
void class_flops(int X, int Y, double &f_prim, double &f_digest);
// internal constructor (api.cu): n_orb > 0 builds the two RI pair groups
int basis_create_internal(int nsh, const double *centers, const int32_t *L, const int32_t *nprim, const double *exps,
                          const double *coefs, int n_orb, qlb_basis **out);
void base_class_params(qlb_basis *b, ClassParams &prm);
// sum over the ranks of the engine's communicator, in place (api.cu; NCCL)
int allreduce_sum_device(double *d, size_t n, cudaStream_t s);
// one-electron kernels (one_electron.cu)
int one_electron_device(qlb_basis *b, int which, int natoms, const double *d_xyzZ, double *d_out, cudaStream_t s);
}  // namespace qlb
