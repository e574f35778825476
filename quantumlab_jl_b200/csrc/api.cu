// api.cu -- the C ABI of libqlb200.so (see include/qlb200.h) and the host-side orchestration of the Fock build:
// pair lists, class bins, screening prefixes, kernel launches, statistics.  No CPU compute path exists here.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "qlb_internal.h"

namespace qlb {

static thread_local std::string g_err;
// One Engine per device.  The engine the calling thread works on follows the handle it passes in (qlb_basis carries its
// engine); qlb_init(device) selects the engine that handle-less calls (qlb_basis_create, qlb_comm_init, ...) use.
constexpr int MAX_DEVICES = 16;
static Engine g_engines[MAX_DEVICES];
static thread_local Engine *g_cur = &g_engines[0];
#define g_engine (*g_cur)
Engine &engine() { return *g_cur; }
void use_engine(Engine *e) {
  if (e && e != g_cur) {
    g_cur = e;
    cudaSetDevice(e->device);
  }
}
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what) {
  set_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
  return QLB_ERR_CUDA;
}

// ------------------------------------------------------------------ Boys table (host, long double series)
static void make_boys_table(std::vector<double> &tab) {
  // layout [L][row][BOYS_ROWLEN]: F_{L+k}(T_i)/k! (k = 0..7), exp(-T_i), pad   (md_core.cuh)
  tab.assign((size_t)(BOYS_LMAX + 1) * BOYS_TABLE_DOUBLES, 0.0);
  const int mtop = BOYS_LMAX + 7;
  std::vector<long double> F(mtop + 1);
  for (int i = 0; i < BOYS_NROW; ++i) {
    const long double T = (long double)i * (long double)BOYS_DT;
    // F_mtop(T) = exp(-T) sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1)), then downward recursion
    long double term = 1.0L / (2 * mtop + 1), sum = term;
    for (int k = 1; k < 2000; ++k) {
      term *= 2.0L * T / (2 * mtop + 2 * k + 1);
      sum += term;
      if (term < 1e-22L * sum) break;
    }
    const long double ex = expl(-T);
    F[mtop] = ex * sum;
    for (int m = mtop; m > 0; --m) F[m - 1] = (2.0L * T * F[m] + ex) / (2 * m - 1);
    for (int L = 0; L <= BOYS_LMAX; ++L) {
      double *row = tab.data() + (size_t)L * BOYS_TABLE_DOUBLES + (size_t)i * BOYS_ROWLEN;
      long double fact = 1.0L;
      for (int k = 0; k < 8; ++k) {
        row[k] = (double)(F[L + k] / fact);
        fact *= (long double)(k + 1);
      }
      row[8] = (double)ex;
    }
  }
}

// ------------------------------------------------------------------ small kernels
// K1: primitive-pair records, one warp per shell pair, lanes stride the pair's KEPT primitive pairs (coalesced SoA writes).
// kept[kept_off[w] + t] = ia * nb + ib of the t-th kept primitive pair (host: prune_primitive_pairs)
__global__ void pair_prim_kernel(int npairs, const int *__restrict__ pair_a, const int *__restrict__ pair_b,
                                 const int *__restrict__ pair_poff, const int *__restrict__ kept_off, const int *__restrict__ kept,
                                 const double *__restrict__ xyz, const int *__restrict__ poff, const double *__restrict__ exps,
                                 const double *__restrict__ coefs, double2 *rec) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= npairs) return;
  const int a = pair_a[w], b = pair_b[w];
  const int oa = poff[a], ob = poff[b], nb = poff[b + 1] - ob;
  const double Ax = xyz[3 * a], Ay = xyz[3 * a + 1], Az = xyz[3 * a + 2];
  const double Bx = xyz[3 * b], By = xyz[3 * b + 1], Bz = xyz[3 * b + 2];
  const double ab2 = (Ax - Bx) * (Ax - Bx) + (Ay - By) * (Ay - By) + (Az - Bz) * (Az - Bz);
  const size_t base = (size_t)pair_poff[w];  // tile layout: primitive t, field f at base + (3 t + f) * 32
  const int k0 = kept_off[w], nk = kept_off[w + 1] - k0;
  for (int t = lane; t < nk; t += 32) {
    const int code = kept[k0 + t];
    const int ia = code / nb, ib = code - ia * nb;
    const double al = exps[oa + ia], be = exps[ob + ib];
    const double p = al + be, ip = 1.0 / p;
    double2 *r = rec + base + (size_t)(3 * TILE_KET) * t;
    r[0] = make_double2(p, (al * Ax + be * Bx) * ip);
    r[TILE_KET] = make_double2((al * Ay + be * By) * ip, (al * Az + be * Bz) * ip);
    r[2 * TILE_KET] = make_double2(coefs[oa + ia] * coefs[ob + ib] * exp(-al * be * ip * ab2) * ip, 0.5 * ip);
  }
}

// K3a: dl[i + nsh*j] = qlog(max |P| over shell block (i,j)); dl[nsh*nsh] = global max
__global__ void density_block_max_kernel(int nsh, int N, const int *__restrict__ boff, const double *__restrict__ P, int *dl) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int q = -32768;
  if (idx < nsh * nsh) {
    const int i = idx % nsh, j = idx / nsh;
    double m = 0.0;
    for (int c = boff[j]; c < boff[j + 1]; ++c)
      for (int r = boff[i]; r < boff[i + 1]; ++r) m = fmax(m, fabs(P[r + (size_t)N * c]));
    q = qlog_exact(m);
    dl[idx] = q;
  }
  q = __reduce_max_sync(0xffffffffu, q);
  if ((threadIdx.x & 31) == 0) atomicMax(dl + nsh * nsh, q);
}

// per build: the density bound of every pair's own shell block, in pair order (read coalesced by the screening walk)
__global__ void pair_dl_kernel(int npairs, const int *__restrict__ pair_a, const int *__restrict__ pair_b, int nsh,
                               const int *__restrict__ dl, int *pair_dl) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < npairs) pair_dl[g] = dl[pair_b[g] + nsh * pair_a[g]];
}

__global__ void symmetrize_kernel(int N, double *A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < N && j < i) {
    const double v = 0.5 * (A[i + (size_t)N * j] + A[j + (size_t)N * i]);
    A[i + (size_t)N * j] = v;
    A[j + (size_t)N * i] = v;
  }
}

// K6: F = H + 2J - K  (SpecialMatricesModule.jl:289)
__global__ void fock_kernel(size_t n, const double *__restrict__ H, const double *__restrict__ J,
                            const double *__restrict__ K, double *F) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) F[i] = H[i] + 2.0 * J[i] - K[i];
}

__global__ void dfma_probe_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

static int require_ready() {
  if (!g_engine.ready) {
    set_error("qlb_init has not been called (or failed): no CUDA device, no compute -- there is no CPU fallback");
    return QLB_ERR_STATE;
  }
  return QLB_OK;
}

// entry points that receive a handle work on the handle's engine (device)
static int enter(const qlb_basis *b) {
  if (b && b->eng) use_engine(b->eng);
  return require_ready();
}

template <class T>
static int upload(T **dptr, const std::vector<T> &h) {
  QLB_CUDA(cudaMalloc((void **)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) QLB_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return QLB_OK;
}

static void fill_pp(const qlb_basis *b, PrimPairs &pp) {
  pp.rec = b->d_pp;
}

static int run_pair_prims(qlb_basis *b) {
  Engine &e = g_engine;
  const int np = (int)b->npairs;
  if (np == 0) return QLB_OK;
  const int threads = 256, blocks = (int)(((int64_t)np * 32 + threads - 1) / threads);
  // (re)upload the kept-primitive lists of the current pair order
  if (b->d_kept_off) cudaFree(b->d_kept_off);
  if (b->d_kept) cudaFree(b->d_kept);
  b->d_kept_off = b->d_kept = nullptr;
  if (int rc = upload(&b->d_kept_off, b->h_kept_off)) return rc;
  if (int rc = upload(&b->d_kept, b->h_kept)) return rc;
  pair_prim_kernel<<<blocks, threads, 0, e.stream>>>(np, b->d_pair_a, b->d_pair_b, b->d_pair_poff, b->d_kept_off, b->d_kept, b->d_xyz,
                                                     b->d_poff, b->d_exps, b->d_coefs, b->d_pp);
  QLB_CUDA(cudaGetLastError());
  return QLB_OK;
}

// Tile layout of the primitive-pair records: within each class (and RI group) the pairs are cut into tiles of 32
// consecutive pairs (= the 32 ket lanes of a warp); a tile with Kmax primitive pairs in its longest pair owns
// 3*32*Kmax double2 slots, ordered [primitive][field][lane].  h_pair_poff[i] = double2 index of (primitive 0, field 0).
// Primitive-pair pruning (new functionality; the reference evaluates every primitive quartet): a primitive pair whose
// prefactor |c_a c_b exp(-a b/(a+b) |AB|^2) / (a+b)| is below PRIM_PAIR_EPS cannot change any integral by more than
// ~1e-19 (2 pi^2.5 |k_ab| |k_cd| / sqrt(p+q) x O(10) Hermite factors), eight orders below the 1e-10 parity bar; on
// (H2O)32 / 6-31G* about half of the primitive pairs of the surviving shell pairs go.  The pairs that stay are listed
// largest prefactor first; every shell pair keeps at least one.
constexpr double PRIM_PAIR_EPS = 1e-22;
static void prune_primitive_pairs(qlb_basis *b) {
  b->h_kept_off.assign(b->npairs + 1, 0);
  b->h_kept.clear();
  std::vector<std::pair<double, int>> cand;
  for (int64_t i = 0; i < b->npairs; ++i) {
    const int a = b->h_pair_a[i], c = b->h_pair_b[i];
    const int oa = b->h_poff[a], oc = b->h_poff[c], na = b->h_nprim[a], nc = b->h_nprim[c];
    const double *A = &b->h_xyz[3 * a], *C = &b->h_xyz[3 * c];
    const double ab2 = (A[0] - C[0]) * (A[0] - C[0]) + (A[1] - C[1]) * (A[1] - C[1]) + (A[2] - C[2]) * (A[2] - C[2]);
    cand.clear();
    for (int ia = 0; ia < na; ++ia)
      for (int ic = 0; ic < nc; ++ic) {
        const double al = b->h_exps[oa + ia], be = b->h_exps[oc + ic], p = al + be;
        const double k = fabs(b->h_coefs[oa + ia] * b->h_coefs[oc + ic]) * exp(-al * be / p * ab2) / p;
        cand.push_back({k, ia * nc + ic});
      }
    std::stable_sort(cand.begin(), cand.end(), [](const std::pair<double, int> &x, const std::pair<double, int> &y) { return x.first > y.first; });
    int keep = 0;
    for (const auto &cd : cand)
      if (keep == 0 || cd.first >= PRIM_PAIR_EPS) {
        b->h_kept.push_back(cd.second);
        ++keep;
      }
    b->h_kept_off[i + 1] = (int)b->h_kept.size();
  }
}

static void compute_pair_offsets(qlb_basis *b) {
  b->h_pair_poff.assign(b->npairs + 1, 0);
  b->h_pair_np.assign(std::max<int64_t>(b->npairs, 1), 0);
  prune_primitive_pairs(b);
  for (int64_t i = 0; i < b->npairs; ++i) b->h_pair_np[i] = b->h_kept_off[i + 1] - b->h_kept_off[i];
  int64_t slot = 0;  // in double2 units
  auto layout = [&](int lo, int hi) {
    for (int t0 = lo; t0 < hi; t0 += TILE_KET) {
      int kmax = 0;
      for (int i = t0; i < std::min(hi, t0 + TILE_KET); ++i) kmax = std::max(kmax, b->h_pair_np[i]);
      for (int i = t0; i < std::min(hi, t0 + TILE_KET); ++i) b->h_pair_poff[i] = (int)(slot + (i - t0));
      slot += (int64_t)3 * TILE_KET * kmax;
    }
  };
  for (int c = 0; c < NCLASS; ++c) layout(b->class_off[c], b->class_off[c + 1]);
  for (int c = 0; c < NCLASS_AUX; ++c) layout(b->ri_class_off[c], b->ri_class_off[c + 1]);
  b->h_pair_poff[b->npairs] = (int)std::min<int64_t>(slot, 0x7fffffff);
  b->nprimpairs = slot;  // double2 slots of the record buffer
}

}  // namespace qlb

using namespace qlb;

// =================================================================== lifetime
extern "C" int qlb_init(int device) {
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device available (") + cudaGetErrorString(err) +
              "): libqlb200 has no CPU fallback");
    return QLB_ERR_CUDA;
  }
  if (device < 0) QLB_CUDA(cudaGetDevice(&device));
  if (device >= ndev || device >= MAX_DEVICES) {
    set_error("device index out of range");
    return QLB_ERR_ARG;
  }
  g_cur = &g_engines[device];
  Engine &e = g_engine;
  QLB_CUDA(cudaSetDevice(device));
  if (e.ready) return QLB_OK;
  cudaDeviceProp prop;
  QLB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error(std::string("device ") + prop.name + " is not sm_100-class; libqlb200 ships sm_100a code only");
    return QLB_ERR_CUDA;
  }
  e.device = device;
  e.sm_count = prop.multiProcessorCount;
  QLB_CUDA(cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking));
  for (int i = 0; i < Engine::NSIDE; ++i) {
    QLB_CUDA(cudaStreamCreateWithFlags(&e.side_stream[i], cudaStreamNonBlocking));
    QLB_CUDA(cudaEventCreateWithFlags(&e.ev_join[i], cudaEventDisableTiming));
  }
  QLB_CUDA(cudaEventCreateWithFlags(&e.ev_fork, cudaEventDisableTiming));
  QLB_CUDA(cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking));
  QLB_CUDA(cudaEventCreateWithFlags(&e.ev_copy, cudaEventDisableTiming));
  std::vector<double> tab;
  make_boys_table(tab);
  QLB_CUDA(cudaMalloc((void **)&e.d_boys, tab.size() * sizeof(double)));
  QLB_CUDA(cudaMemcpy(e.d_boys, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  // the rolled kernels (tensor / RI / Schwarz use) keep their intermediates in local memory: the largest frame is the
  // (dd|dd) Schwarz / rolled class kernel's (ptxas -v: 15.3 KB)
  QLB_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, 16 * 1024));
  if (const char *v = getenv("QLB_YS_TARGET")) e.ys_target = std::max(1, atoi(v));
  if (const char *v = getenv("QLB_YS_TILES")) e.ys_tiles = std::max(1, atoi(v));
  if (const char *v = getenv("QLB_LAUNCH_ORDER")) e.launch_order = atoi(v);
  if (const char *v = getenv("QLB_NSIDE")) e.nside = std::max(0, std::min(atoi(v), (int)Engine::NSIDE));
  if (const char *v = getenv("QLB_DLSMEM_MIN_TILES")) e.dlsmem_min_tiles = std::max(0, atoi(v));
  e.ysplit_full = getenv("QLB_YSPLIT_FULL") != nullptr;
  if (const char *v = getenv("QLB_BIN_W")) e.bin_width = std::max(1, atoi(v));
  e.no_schedule = getenv("QLB_NO_SCHEDULE") != nullptr;
  e.no_dl_smem = getenv("QLB_NO_DLSMEM") != nullptr;
  e.ready = true;
  return QLB_OK;
}

extern "C" int qlb_finalize(void) {
  Engine *keep = g_cur;
  for (int d = 0; d < MAX_DEVICES; ++d) {
    Engine &e = g_engines[d];
    if (!e.ready) continue;
    g_cur = &e;
    cudaSetDevice(e.device);
    qlb_comm_destroy();
    if (e.d_boys) cudaFree(e.d_boys);
    if (e.stream) cudaStreamDestroy(e.stream);
    for (int i = 0; i < Engine::NSIDE; ++i) {
      if (e.side_stream[i]) cudaStreamDestroy(e.side_stream[i]);
      if (e.ev_join[i]) cudaEventDestroy(e.ev_join[i]);
    }
    if (e.ev_fork) cudaEventDestroy(e.ev_fork);
    if (e.copy_stream) cudaStreamDestroy(e.copy_stream);
    if (e.ev_copy) cudaEventDestroy(e.ev_copy);
    e.copy_stream = nullptr;
    e.ev_copy = nullptr;
    e = Engine();
  }
  g_cur = keep;
  return QLB_OK;
}

extern "C" const char *qlb_last_error(void) { return g_err.c_str(); }

extern "C" int qlb_device_info(char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor) {
  if (int rc = require_ready()) return rc;
  cudaDeviceProp prop;
  QLB_CUDA(cudaGetDeviceProperties(&prop, g_engine.device));
  if (name && name_len > 0) {
    strncpy(name, prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return QLB_OK;
}

// the quantised log used by the screening rule (host evaluation of the same inline function the kernels use)
extern "C" int qlb_qlog(double x) { return qlog_exact(x); }

extern "C" int qlb_set_profiling(int on) {
  g_engine.profiling = on != 0;
  return QLB_OK;
}

extern "C" int qlb_fp64_peak_probe(double *tflops) {
  if (int rc = require_ready()) return rc;
  Engine &e = g_engine;
  QLB_CUDA(cudaSetDevice(e.device));
  const int blocks = e.sm_count * 8, threads = 256, iters = 1 << 15;
  double *d = nullptr;
  QLB_CUDA(cudaMalloc((void **)&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  dfma_probe_kernel<<<blocks, threads, 0, e.stream>>>(d, 1024);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, e.stream);
    dfma_probe_kernel<<<blocks, threads, 0, e.stream>>>(d, iters);
    cudaEventRecord(b, e.stream);
    QLB_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    best = std::max(best, fl / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  if (tflops) *tflops = best;
  return QLB_OK;
}

// =================================================================== basis
extern "C" int qlb_basis_destroy(qlb_basis *b) {
  if (!b) return QLB_OK;
  if (b->eng) use_engine(b->eng);
  if (g_engine.ready) cudaSetDevice(g_engine.device);
  void *ptrs[] = {b->d_xyz, b->d_exps, b->d_coefs, b->d_poff, b->d_boff, b->d_L, b->d_pair_a, b->d_pair_b,
                  b->d_pair_poff, b->d_pair_np, b->d_pair_ql, b->d_pair_qls, b->d_pair_Q, b->d_pp, b->d_pair_info, b->d_pair_dl, b->d_kept_off, b->d_kept,
                  b->d_P, b->d_JK, b->d_H, b->d_dl, b->d_counters, b->d_ms_local};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (auto &ev : b->ev)
    if (ev) cudaEventDestroy(ev);
  if (b->ev_gcounts) cudaEventDestroy(b->ev_gcounts);
  if (b->ev_dlmax) cudaEventDestroy(b->ev_dlmax);
  if (b->h_gcounts) cudaFreeHost(b->h_gcounts);
  if (b->h_dlmax) cudaFreeHost(b->h_dlmax);
  delete b;
  return QLB_OK;
}

static int basis_build_pairs(qlb_basis *b) {
  Engine &e = g_engine;
  const int nsh = b->nsh;
  // enumerate unique pairs, canonical orientation la >= lb, bucketed by class in (i,j) order
  std::vector<int> pa[NCLASS + NCLASS_AUX], pb[NCLASS + NCLASS_AUX];
  const int nreg = b->n_orb > 0 ? b->n_orb : nsh;
  for (int i = 0; i < nreg; ++i)
    for (int j = 0; j <= i; ++j) {
      int a = i, c = j;
      if (b->h_L[a] < b->h_L[c]) std::swap(a, c);
      const int cls = class_of(b->h_L[a], b->h_L[c]);
      pa[cls].push_back(a);
      pb[cls].push_back(c);
    }
  if (b->n_orb > 0)  // RI: (auxiliary shell, unit s function) pairs form the second group
    for (int k = b->n_orb; k < nsh - 1; ++k) {
      const int cls = NCLASS + class_of(b->h_L[k], 0);
      pa[cls].push_back(k);
      pb[cls].push_back(nsh - 1);
    }
  b->h_pair_a.clear();
  b->h_pair_b.clear();
  b->class_off[0] = 0;
  for (int c = 0; c < NCLASS; ++c) {
    b->h_pair_a.insert(b->h_pair_a.end(), pa[c].begin(), pa[c].end());
    b->h_pair_b.insert(b->h_pair_b.end(), pb[c].begin(), pb[c].end());
    b->class_off[c + 1] = (int)b->h_pair_a.size();
  }
  b->ri_class_off[0] = (int)b->h_pair_a.size();
  for (int c = 0; c < NCLASS_AUX; ++c) {
    b->h_pair_a.insert(b->h_pair_a.end(), pa[NCLASS + c].begin(), pa[NCLASS + c].end());
    b->h_pair_b.insert(b->h_pair_b.end(), pb[NCLASS + c].begin(), pb[NCLASS + c].end());
    b->ri_class_off[c + 1] = (int)b->h_pair_a.size();
  }
  b->npairs = (int64_t)b->h_pair_a.size();
  compute_pair_offsets(b);
  if (b->nprimpairs > 0x7fffffff) {
    set_error("too many primitive pairs for 32-bit offsets");
    return QLB_ERR_ARG;
  }
  if (int rc = upload(&b->d_pair_a, b->h_pair_a)) return rc;
  if (int rc = upload(&b->d_pair_b, b->h_pair_b)) return rc;
  if (int rc = upload(&b->d_pair_poff, b->h_pair_poff)) return rc;
  QLB_CUDA(cudaMalloc((void **)&b->d_pp, std::max<int64_t>(b->nprimpairs, 1) * sizeof(double2)));
  if (int rc = upload(&b->d_pair_np, b->h_pair_np)) return rc;
  QLB_CUDA(cudaMalloc((void **)&b->d_pair_Q, std::max<int64_t>(b->npairs, 1) * sizeof(double)));
  QLB_CUDA(cudaMalloc((void **)&b->d_pair_ql, std::max<int64_t>(b->npairs, 1) * sizeof(int)));
  QLB_CUDA(cudaMalloc((void **)&b->d_pair_qls, std::max<int64_t>(b->npairs, 1) * sizeof(int)));
  if (int rc = run_pair_prims(b)) return rc;
  // K2: Schwarz bounds per pair (the (aux, unit) pairs of an RI basis keep bound 0: they are never screened)
  QLB_CUDA(cudaMemsetAsync(b->d_pair_Q, 0, std::max<int64_t>(b->npairs, 1) * sizeof(double), e.stream));
  SchwarzParams sp;
  sp.sh_xyz = b->d_xyz; sp.pair_a = b->d_pair_a; sp.pair_b = b->d_pair_b; sp.pair_poff = b->d_pair_poff; sp.pair_np = b->d_pair_np;
  fill_pp(b, sp.pp);
  sp.boys = e.d_boys; sp.Q = b->d_pair_Q;
  for (int c = 0; c < NCLASS; ++c) {
    sp.off = b->class_off[c];
    sp.n = b->class_off[c + 1] - b->class_off[c];
    if (sp.n == 0) continue;
    if (int ce = launch_schwarz_kernel(c, (unsigned)((sp.n + 63) / 64), e.stream, sp)) return cuda_fail((cudaError_t)ce, "schwarz_kernel");
  }
  b->h_pair_Q.resize(b->npairs);
  QLB_CUDA(cudaMemcpyAsync(b->h_pair_Q.data(), b->d_pair_Q, b->npairs * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  // K3 (host part): order the pairs of each class by quantised Schwarz bound (descending), ties by shell indices.  The
  // walks of the kernels end where the running maximum of the bounds (pair_qls) fails.  Engine::bin_width > 1 coarsens
  // the bound into bins and orders by shell index inside a bin (index-local tiles: fewer cache lines per gather) -- an
  // experiment that lost on every class (mixed survivors inside a tile cost more than the gathers save), kept as a knob
  // (QLB_BIN_W).  The screening DECISION per quartet does not depend on the order.
  std::vector<int> perm(b->npairs);
  std::iota(perm.begin(), perm.end(), 0);
  std::vector<int> ql0(b->npairs);
  for (int64_t i = 0; i < b->npairs; ++i) ql0[i] = qlog_exact(b->h_pair_Q[i]);
  const int binw = std::max(1, g_engine.bin_width);
  b->seg_off.clear();
  b->seg_cls.clear();
  for (int c = 0; c < NCLASS; ++c) {
    int qmax = -32768;
    for (int i = b->class_off[c]; i < b->class_off[c + 1]; ++i) qmax = std::max(qmax, ql0[i]);
    auto bin = [&](int x) { return ql0[x] <= -32768 ? (1 << 28) : (qmax - ql0[x]) / binw; };
    std::stable_sort(perm.begin() + b->class_off[c], perm.begin() + b->class_off[c + 1], [&](int x, int y) {
      const int bx = bin(x), by = bin(y);
      if (bx != by) return bx < by;
      if (b->h_pair_a[x] != b->h_pair_a[y]) return b->h_pair_a[x] < b->h_pair_a[y];
      return b->h_pair_b[x] < b->h_pair_b[y];
    });
    if (b->class_off[c + 1] > b->class_off[c]) {
      b->seg_off.push_back(b->class_off[c]);
      b->seg_cls.push_back(c);
    }
  }
  b->seg_off.push_back(b->class_off[NCLASS]);
  std::vector<int> na(b->npairs), nb(b->npairs);
  std::vector<double> nq(b->npairs);
  b->h_pair_ql.resize(b->npairs);
  b->h_pair_qls.resize(b->npairs);
  b->qlmax = -32768;
  for (int64_t i = 0; i < b->npairs; ++i) {
    na[i] = b->h_pair_a[perm[i]];
    nb[i] = b->h_pair_b[perm[i]];
    nq[i] = b->h_pair_Q[perm[i]];
    b->h_pair_ql[i] = qlog_exact(nq[i]);
    b->qlmax = std::max(b->qlmax, b->h_pair_ql[i]);
  }
  // pair_qls: running maximum from the END of each class (and RI group): bounds this pair and every later one
  auto suffix_max = [&](int lo, int hi) {
    int m = -32768;
    for (int i = hi - 1; i >= lo; --i) {
      m = std::max(m, b->h_pair_ql[i]);
      b->h_pair_qls[i] = m;
    }
  };
  for (int c = 0; c < NCLASS; ++c) suffix_max(b->class_off[c], b->class_off[c + 1]);
  for (int c = 0; c < NCLASS_AUX; ++c) suffix_max(b->ri_class_off[c], b->ri_class_off[c + 1]);
  b->h_pair_a.swap(na);
  b->h_pair_b.swap(nb);
  b->h_pair_Q.swap(nq);
  compute_pair_offsets(b);
  if (b->nprimpairs > 0x7fffffff) {
    set_error("too many primitive pairs for 32-bit offsets");
    return QLB_ERR_ARG;
  }
  cudaFree(b->d_pp);  // the tile layout of the sorted order needs a different number of slots
  b->d_pp = nullptr;
  QLB_CUDA(cudaMalloc((void **)&b->d_pp, std::max<int64_t>(b->nprimpairs, 1) * sizeof(double2)));
  QLB_CUDA(cudaMemcpy(b->d_pair_a, b->h_pair_a.data(), b->npairs * sizeof(int), cudaMemcpyHostToDevice));
  QLB_CUDA(cudaMemcpy(b->d_pair_b, b->h_pair_b.data(), b->npairs * sizeof(int), cudaMemcpyHostToDevice));
  QLB_CUDA(cudaMemcpy(b->d_pair_poff, b->h_pair_poff.data(), (b->npairs + 1) * sizeof(int), cudaMemcpyHostToDevice));
  QLB_CUDA(cudaMemcpy(b->d_pair_np, b->h_pair_np.data(), b->npairs * sizeof(int), cudaMemcpyHostToDevice));
  QLB_CUDA(cudaMemcpy(b->d_pair_ql, b->h_pair_ql.data(), b->npairs * sizeof(int), cudaMemcpyHostToDevice));
  QLB_CUDA(cudaMemcpy(b->d_pair_qls, b->h_pair_qls.data(), b->npairs * sizeof(int), cudaMemcpyHostToDevice));
  {
    // packed per-pair records (eri_kernels.cuh: load_pair_info)
    std::vector<double2> info((size_t)PAIR_INFO_BLOCK * std::max<int64_t>((b->npairs + 31) / 32, 1));
    for (int64_t i = 0; i < b->npairs; ++i) {
      const int pa_ = b->h_pair_a[i], pb_ = b->h_pair_b[i];
      const double *A = &b->h_xyz[3 * pa_], *B = &b->h_xyz[3 * pb_];
      double2 *r = &info[(size_t)(i >> 5) * PAIR_INFO_BLOCK + (i & 31)];
      r[0] = make_double2(A[0], A[1]);
      r[32] = make_double2(A[2], B[0]);
      r[64] = make_double2(B[1], B[2]);
      const int i3[4] = {pa_, pb_, b->h_pair_poff[i], b->h_pair_np[i]};
      const int i4[4] = {b->h_boff[pa_], b->h_boff[pb_], b->h_pair_ql[i], 0};
      memcpy(&r[96], i3, 16);
      memcpy(&r[128], i4, 16);
    }
    if (b->d_pair_info) cudaFree(b->d_pair_info);
    b->d_pair_info = nullptr;
    QLB_CUDA(cudaMalloc((void **)&b->d_pair_info, info.size() * sizeof(double2)));
    QLB_CUDA(cudaMemcpy(b->d_pair_info, info.data(), info.size() * sizeof(double2), cudaMemcpyHostToDevice));
    if (!b->d_pair_dl) {
      QLB_CUDA(cudaMalloc((void **)&b->d_pair_dl, std::max<int64_t>(b->npairs, 1) * sizeof(int)));
      QLB_CUDA(cudaMemset(b->d_pair_dl, 0, std::max<int64_t>(b->npairs, 1) * sizeof(int)));
    }
  }
  QLB_CUDA(cudaMemcpy(b->d_pair_Q, b->h_pair_Q.data(), b->npairs * sizeof(double), cudaMemcpyHostToDevice));
  if (int rc = run_pair_prims(b)) return rc;
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  return QLB_OK;
}

extern "C" int qlb_basis_create(int nsh, const double *centers, const int32_t *L, const int32_t *nprim,
                                const double *exps, const double *coefs, qlb_basis **out) {
  return qlb::basis_create_internal(nsh, centers, L, nprim, exps, coefs, 0, out);
}

int qlb::basis_create_internal(int nsh, const double *centers, const int32_t *L, const int32_t *nprim, const double *exps,
                               const double *coefs, int n_orb, qlb_basis **out) {
  if (int rc = require_ready()) return rc;
  if (!out || nsh <= 0 || !centers || !L || !nprim || !exps || !coefs) {
    set_error("qlb_basis_create: null pointer or nsh <= 0");
    return QLB_ERR_ARG;
  }
  QLB_CUDA(cudaSetDevice(g_engine.device));
  qlb_basis *b = new qlb_basis();
  b->eng = g_cur;
  b->nsh = nsh;
  b->n_orb = n_orb;
  b->h_xyz.assign(centers, centers + 3 * (size_t)nsh);
  b->h_L.assign(L, L + nsh);
  b->h_nprim.assign(nprim, nprim + nsh);
  b->h_poff.resize(nsh + 1);
  b->h_boff.resize(nsh + 1);
  b->h_poff[0] = b->h_boff[0] = 0;
  for (int i = 0; i < nsh; ++i) {
    // orbital shells: s, p, d; auxiliary shells of an RI basis (indices >= n_orb): up to f
    const int lmax = (n_orb > 0 && i >= n_orb) ? QLB_MAX_L_AUX : QLB_MAX_L;
    if (L[i] < 0 || L[i] > lmax) {
      // the reference performs the analogous check against MAX_AM on the Julia side (LibInt2Module.jl:79-83)
      set_error("shell " + std::to_string(i) + ": angular momentum " + std::to_string(L[i]) + " outside 0.." + std::to_string(lmax) +
                " (QLB_MAX_L for orbital shells, QLB_MAX_L_AUX for auxiliary shells)");
      delete b;
      return QLB_ERR_ARG;
    }
    if (nprim[i] <= 0) {
      set_error("qlb_basis_create: shell without primitives");
      delete b;
      return QLB_ERR_ARG;
    }
    b->h_poff[i + 1] = b->h_poff[i] + nprim[i];
    b->h_boff[i + 1] = b->h_boff[i] + ncart(L[i]);
  }
  b->nprim_total = b->h_poff[nsh];
  b->nbf = b->h_boff[nsh];
  if (n_orb == 0)  // public entry: every primitive needs a positive, finite exponent (the RI unit function is internal)
    for (int i = 0; i < b->nprim_total; ++i)
      if (!(exps[i] > 0.0) || !std::isfinite(exps[i]) || !std::isfinite(coefs[i])) {
        set_error("qlb_basis_create: exponents must be positive and finite, coefficients finite");
        delete b;
        return QLB_ERR_ARG;
      }
  b->h_exps.assign(exps, exps + b->nprim_total);
  b->h_coefs.assign(coefs, coefs + b->nprim_total);
  int rc = QLB_OK;
  if ((rc = upload(&b->d_xyz, b->h_xyz)) || (rc = upload(&b->d_exps, b->h_exps)) || (rc = upload(&b->d_coefs, b->h_coefs)) ||
      (rc = upload(&b->d_poff, b->h_poff)) || (rc = upload(&b->d_boff, b->h_boff)) || (rc = upload(&b->d_L, b->h_L))) {
    qlb_basis_destroy(b);
    return rc;
  }
  const size_t N2 = (size_t)b->nbf * b->nbf;
  cudaError_t ce;
  if ((ce = cudaMalloc((void **)&b->d_P, N2 * sizeof(double))) != cudaSuccess ||
      (ce = cudaMalloc((void **)&b->d_JK, (2 * N2 + QLB_JK_TAIL) * sizeof(double))) != cudaSuccess ||
      (ce = cudaMalloc((void **)&b->d_H, N2 * sizeof(double))) != cudaSuccess ||
      (ce = cudaMalloc((void **)&b->d_dl, ((size_t)nsh * nsh + 1) * sizeof(int))) != cudaSuccess ||
      (ce = cudaMalloc((void **)&b->d_counters, 2 * NQCLASS * sizeof(unsigned long long))) != cudaSuccess) {
    qlb_basis_destroy(b);
    return cuda_fail(ce, "cudaMalloc(scratch)");
  }
  for (auto &ev : b->ev) cudaEventCreate(&ev);
  cudaEventCreateWithFlags(&b->ev_gcounts, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&b->ev_dlmax, cudaEventDisableTiming);
  if ((ce = cudaMalloc((void **)&b->d_ms_local, NQCLASS * sizeof(double))) != cudaSuccess ||
      (ce = cudaMallocHost((void **)&b->h_dlmax, sizeof(int))) != cudaSuccess ||
      (ce = cudaMallocHost((void **)&b->h_gcounts, 3 * NQCLASS * sizeof(double))) != cudaSuccess) {
    qlb_basis_destroy(b);
    return cuda_fail(ce, "cudaMalloc(gcounts)");
  }
  if ((rc = basis_build_pairs(b))) {
    qlb_basis_destroy(b);
    return rc;
  }
  *out = b;
  return QLB_OK;
}

extern "C" int qlb_basis_nbf(const qlb_basis *b, int *nbf) {
  if (!b || !nbf) return QLB_ERR_ARG;
  *nbf = b->nbf;
  return QLB_OK;
}
extern "C" int qlb_basis_npairs(const qlb_basis *b, int64_t *npairs) {
  if (!b || !npairs) return QLB_ERR_ARG;
  *npairs = b->npairs;
  return QLB_OK;
}

extern "C" int qlb_schwarz(qlb_basis *b, double *Q) {
  if (int rc = enter(b)) return rc;
  if (!b || !Q) return QLB_ERR_ARG;
  const int n = b->nsh;
  for (int64_t i = 0; i < b->npairs; ++i) {
    const int a = b->h_pair_a[i], c = b->h_pair_b[i];
    Q[a + (size_t)n * c] = Q[c + (size_t)n * a] = b->h_pair_Q[i];
  }
  return QLB_OK;
}

// =================================================================== J/K
static void base_params(qlb_basis *b, ClassParams &prm);
void qlb::base_class_params(qlb_basis *b, ClassParams &prm) { base_params(b, prm); }

static void base_params(qlb_basis *b, ClassParams &prm) {
  memset(&prm, 0, sizeof(prm));
  prm.sh_xyz = b->d_xyz; prm.sh_boff = b->d_boff; prm.nsh = b->nsh; prm.nbf = b->nbf;
  prm.pair_a = b->d_pair_a; prm.pair_b = b->d_pair_b; prm.pair_poff = b->d_pair_poff; prm.pair_np = b->d_pair_np; prm.pair_ql = b->d_pair_ql; prm.pair_qls = b->d_pair_qls;
  prm.pair_dl = b->d_pair_dl; prm.pair_info = b->d_pair_info;
  prm.dl_smem = (b->nsh <= DL_SMEM_MAX_NSH && !g_engine.no_dl_smem) ? 1 : 0;
  fill_pp(b, prm.pp);
  prm.boys = g_engine.d_boys;
  prm.dl = b->d_dl;
  prm.nranks = 1;
}

// Dealing class pieces to ranks.  cost[qc] is the MEASURED cost (ms) of class pair qc (class kernel time of the calibration
// build summed over the ranks, scaled by the class's current work): no compiled-in efficiency table.
//  * A class whose per-rank share is at least SPLIT_MIN_MS is cut into nranks pieces (row subsets p, p+nranks, ... of the
//    bound-sorted rows), one per rank: balanced by construction.  (Measured on 2 GPUs, (H2O)32: dealing whole classes
//    longest-first by their calibration times gave 20.1 ms per build, dealing rows 18.3 ms -- class times measured one
//    kernel at a time do not add up under concurrent streams.)
//  * Smaller classes would become launch/tail-bound slivers: they are cut into pieces of about SPLIT_MIN_MS and the pieces
//    assigned longest-first to the least loaded rank.
// Pure function of (cost, nranks): identical on all ranks.
constexpr double SPLIT_MIN_MS = 0.12;
static bool build_schedule(const double *cost, int rank, int nranks, std::vector<std::pair<int, int>> *mine) {
  double total = 0.0;
  for (int qc = 0; qc < NQCLASS; ++qc) total += std::max(cost[qc], 0.0);
  if (!(total > 0.0)) return false;
  struct Piece { double c; int qc, p, np; };
  std::vector<Piece> pieces;
  std::vector<double> load(nranks, 0.0);
  for (int qc = 0; qc < NQCLASS; ++qc) {
    const double c = std::max(cost[qc], 0.0);
    if (c >= SPLIT_MIN_MS * nranks) {
      // piece p goes to rank (p + qc) mod nranks: the rank that receives row 0 (the heaviest) rotates with the class
      for (int r = 0; r < nranks; ++r) load[r] += c / nranks;
      mine[qc].push_back({(rank + qc) % nranks, nranks});
      continue;
    }
    const int np = std::max(1, std::min((int)floor(c / SPLIT_MIN_MS), nranks));
    for (int p = 0; p < np; ++p) pieces.push_back({c / np, qc, p, np});
  }
  std::stable_sort(pieces.begin(), pieces.end(), [](const Piece &a, const Piece &b) { return a.c > b.c; });
  for (const Piece &pc : pieces) {
    int r = 0;
    for (int k = 1; k < nranks; ++k)
      if (load[k] < load[r]) r = k;
    load[r] += pc.c;
    if (r == rank) mine[pc.qc].push_back({pc.p, pc.np});
  }
  return true;
}

// Exposes the schedule for inspection / tests: for each class pair the number of pieces and the owning rank of each
// piece (owner[qc*nranks + p], -1 beyond npieces[qc]).  Pure host function: no device needed.
extern "C" int qlb_schedule_pieces(const double *cost, int nranks, int32_t *npieces, int32_t *owner) {
  if (!cost || !npieces || !owner || nranks < 1) return QLB_ERR_ARG;
  for (int i = 0; i < NQCLASS; ++i) npieces[i] = 0;
  for (int i = 0; i < NQCLASS * nranks; ++i) owner[i] = -1;
  for (int r = 0; r < nranks; ++r) {
    std::vector<std::pair<int, int>> mine[NQCLASS];
    if (!build_schedule(cost, r, nranks, mine)) return QLB_ERR_ARG;
    for (int qc = 0; qc < NQCLASS; ++qc)
      for (const auto &pc : mine[qc]) {
        if (pc.first < 0 || pc.first >= nranks || owner[qc * nranks + pc.first] != -1) return QLB_ERR_STATE;
        owner[qc * nranks + pc.first] = r;
        npieces[qc] = pc.second;
      }
  }
  return QLB_OK;
}

// work of a class pair in model flops: only used as a RELATIVE scale between two builds of the same class
static double class_work(int qc, double nquart, double nprim) {
  for (int X = 0; X < NCLASS; ++X)
    for (int Y = 0; Y <= X; ++Y)
      if (qclass_of(X, Y) == qc) {
        double fp, fd;
        class_flops(X, Y, fp, fd);
        return fp * nprim + fd * nquart;
      }
  return 0.0;
}

// Test hook: injects the schedule feedback a real multi-rank run would have produced (all-rank counters and class times
// of a calibration build), so that one device can evaluate every piece of the scheduled deal (tests/test_gpu_parity.py).
extern "C" int qlb_debug_set_schedule_feedback(qlb_basis *b, int nranks, const double *counts /*2*NQCLASS*/, const double *class_ms /*NQCLASS*/) {
  if (!b || !counts || !class_ms || nranks < 1) return QLB_ERR_ARG;
  if (int rc = enter(b)) return rc;
  for (int i = 0; i < 2 * NQCLASS; ++i) b->h_gcounts[i] = counts[i];
  for (int qc = 0; qc < NQCLASS; ++qc) {
    b->calib_ms[qc] = class_ms[qc];
    b->calib_work[qc] = class_work(qc, counts[2 * qc], counts[2 * qc + 1]);
  }
  b->calib_nranks = nranks;
  b->gcounts_pending = true;
  b->gcounts_nranks = nranks;
  cudaEventRecord(b->ev_gcounts, g_engine.stream);
  return QLB_OK;
}

// policy of a class pair (class_tu.cu, mode -1): bra pairs per CTA and kernels launched per class-pair launch
static int tile_bra_of(int X, int Y) { return launch_class_kernel(X, Y, -1, 0, nullptr, ClassParams()) & 0xff; }
// CTAs per SM of one class launch (sizes ysplit, the CTAs per bra tile row).  A property of the class KERNEL: light classes
// want many short walks (their batches are cheap, the GPU must be filled), the cooperative high-L classes few long ones
// (fuller batches, the bra-side Hermite coefficients staged once per walk).  Measured per class on two unlike workloads,
// (H2O)32/6-31G* and C40H82/def2-SVP, with scripts/shard_classes.py over 4..512 (gpurun_out r2x): the best value per class is
// the same on both to within one step of the sweep; QLB_YS_TARGET overrides all of them.
static int ys_target_of(int X, int Y, const Engine &e) {
  if (e.ys_target > 0) return e.ys_target;
  int la, lb, lc, ld;
  class_L(X, la, lb);
  class_L(Y, lc, ld);
  const int L = la + lb + lc + ld;
  if (lc + ld == 0 && lb == 0) return 64;                       // (ss|ss) (ps|ss) (ds|ss)
  if (tile_bra_of(X, Y) == 1) {                                 // cooperative flavour
    if (la == 2 && lb == 1 && lc == 1 && ld == 1) return 32;    // (dp|pp)
    if (L == 5 || (la + lb == 3 && lc + ld == 3)) return 16;    // (dd|ps) (dp|dp)
    return 8;                                                   // (dd|pp) (dd|ds) (dd|dp) (dd|dd)
  }
  if (L <= 3 || (la + lb == 3 && lc + ld == 1)) return 32;      // (ps|ps) (pp|ss) (pp|ps) (ds|ps) (dp|ss) (dp|ps)
  return 16;                                                    // (pp|pp) (ds|pp) (ds|ds) (dp|ds) (dd|ss)
}
static int launch_count_of(int X, int Y) { return launch_class_kernel(X, Y, -1, 0, nullptr, ClassParams()) >> 8; }

// number of leading pairs of segment sg (sorted by bound, descending) with ql >= need
static int prefix_count(const qlb_basis *b, int sg, int need) {
  const int *lo = b->h_pair_qls.data() + b->seg_off[sg], *hi = b->h_pair_qls.data() + b->seg_off[sg + 1];
  // pair_qls is non-increasing: first position from which no pair reaches `need`
  return (int)(std::partition_point(lo, hi, [need](int q) { return q >= need; }) - lo);
}

extern "C" int qlb_build_jk_device(qlb_basis *b, const double *dP, double tau, double *dJ, double *dK, int rank,
                                   int nranks, void *stream_, qlb_stats *stats) {
  if (int rc = enter(b)) return rc;
  if (!b || !dP || !dJ || !dK || nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("qlb_build_jk_device: bad argument");
    return QLB_ERR_ARG;
  }
  Engine &e = g_engine;
  QLB_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : e.stream;
  const int nsh = b->nsh, N = b->nbf;
  const size_t N2 = (size_t)N * N;
  const bool screen = tau > 0.0;
  // a sharded build of a basis that has no calibration yet times its class kernels one by one (and so does profiling)
  const bool calibrate = nranks > 1 && b->calib_nranks != nranks && !e.no_schedule;
  const bool prof = (e.profiling && stats) || calibrate;
  int launches = 0;
  if (stats) cudaEventRecord(b->ev[0], s);
  // The density bound (max over shell blocks of qlog|P|) never leaves the device: the kernels read it from dl[nsh*nsh].
  // The host only needs it to SIZE the grids, for which the previous build's value (+ one binary order of margin) does;
  // the kernels walk rows with a grid stride and end on the sorted bounds, so a stale value cannot lose a quartet.
  int dl_guess = 1 << 20, tlog = -32768;
  if (screen) {
    if (b->dlmax_pending) {
      QLB_CUDA(cudaEventSynchronize(b->ev_dlmax));  // recorded a whole build ago: already complete
      // + one binary order of margin; never below qlog(1e-3): a build on a zero / tiny density (ZeroGuess, density
      // differences) must not shrink the NEXT build's grids to nothing
      dl_guess = std::max(*b->h_dlmax + 8, -80);
    }
    const int init = -32768;
    QLB_CUDA(cudaMemcpyAsync(b->d_dl + (size_t)nsh * nsh, &init, sizeof(int), cudaMemcpyHostToDevice, s));
    const int threads = 128, blocks = (nsh * nsh + threads - 1) / threads;
    density_block_max_kernel<<<blocks, threads, 0, s>>>(nsh, N, b->d_boff, dP, b->d_dl);
    pair_dl_kernel<<<(unsigned)((b->npairs + 255) / 256), 256, 0, s>>>((int)b->npairs, b->d_pair_a, b->d_pair_b, nsh, b->d_dl, b->d_pair_dl);
    QLB_CUDA(cudaGetLastError());
    launches += 2;
    QLB_CUDA(cudaMemcpyAsync(b->h_dlmax, b->d_dl + (size_t)nsh * nsh, sizeof(int), cudaMemcpyDeviceToHost, s));
    QLB_CUDA(cudaEventRecord(b->ev_dlmax, s));
    b->dlmax_pending = true;
    tlog = qlog_exact(tau);
  }
  QLB_CUDA(cudaMemsetAsync(dJ, 0, N2 * sizeof(double), s));
  QLB_CUDA(cudaMemsetAsync(dK, 0, N2 * sizeof(double), s));
  QLB_CUDA(cudaMemsetAsync(b->d_counters, 0, 2 * NQCLASS * sizeof(unsigned long long), s));
  const int nseg = (int)b->seg_cls.size();
  // pairs that can survive at all (grid sizing only)
  auto sig_count = [&](int sg, int dlmax) {
    const int n = b->seg_off[sg + 1] - b->seg_off[sg];
    if (!screen) return n;
    const int64_t need = (int64_t)tlog - b->qlmax - dlmax;
    return need < -40000 ? n : prefix_count(b, sg, (int)need);
  };
  ClassParams prm;
  base_params(b, prm);
  prm.tlog = tlog; prm.screen = screen ? 1 : 0;
  prm.rank = rank; prm.nranks = nranks;
  prm.P = dP; prm.J = dJ; prm.K = dK;
  const int dl_smem_base = prm.dl_smem;
  if (stats) cudaEventRecord(b->ev[1], s);
  bool qc_started[NQCLASS] = {false};
  // ---- multi-rank schedule from measured class costs (identical on every rank)
  std::vector<std::pair<int, int>> my_pieces[NQCLASS];  // per class pair: (piece index, number of pieces)
  bool scheduled = false;
  if (nranks > 1 && !calibrate && b->gcounts_pending && b->gcounts_nranks == nranks && !e.no_schedule) {
    QLB_CUDA(cudaEventSynchronize(b->ev_gcounts));  // the previous build's all-reduce: long complete
    if (b->calib_work[0] == 0.0 && b->calib_ms[0] == 0.0) {
      // first build after the calibration build: adopt its all-rank class times and work
      bool any = false;
      for (int qc = 0; qc < NQCLASS; ++qc) {
        b->calib_ms[qc] = b->h_gcounts[2 * NQCLASS + qc];
        b->calib_work[qc] = class_work(qc, b->h_gcounts[2 * qc], b->h_gcounts[2 * qc + 1]);
        any = any || b->calib_ms[qc] > 0.0;
      }
      if (!any) b->calib_nranks = 0;  // nothing was measured: calibrate again
    }
    double cost[NQCLASS];
    for (int qc = 0; qc < NQCLASS; ++qc) {
      const double w = class_work(qc, b->h_gcounts[2 * qc], b->h_gcounts[2 * qc + 1]);
      cost[qc] = b->calib_work[qc] > 0.0 ? b->calib_ms[qc] * (w / b->calib_work[qc]) : (w > 0.0 ? 1e-3 : 0.0);
    }
    scheduled = b->calib_nranks == nranks && build_schedule(cost, rank, nranks, my_pieces);
  }
  bool forked = false;
  int nlaunch_classes = 0;
  // heaviest classes first; within a class pair, every (bra segment, ket segment) combination is one launch
  int cls_seg[NCLASS + 1];  // segment range of each class
  for (int c = 0, sg = 0; c <= NCLASS; ++c) {
    while (sg < nseg && b->seg_cls[sg] < c) ++sg;
    cls_seg[c] = sg;
  }
  // Launch order of the class pairs: the launches are dealt over the build stream and the side streams, and the build
  // ends when the last kernel's last CTA does -- the classes expected to run longest go first.  Expected cost = significant
  // bra pairs x significant ket pairs x flops per primitive quartet of the class (host-side quantities only).
  std::vector<std::pair<int, int>> xy_order;
  {
    std::vector<std::pair<double, std::pair<int, int>>> keyed;
    for (int X = NCLASS - 1; X >= 0; --X)
      for (int Y = X; Y >= 0; --Y) {
        double est = 0.0;
        for (int sx = cls_seg[X]; sx < cls_seg[X + 1]; ++sx)
          for (int sy = cls_seg[Y]; sy < cls_seg[Y + 1]; ++sy) {
            double fp, fd;
            class_flops(X, Y, fp, fd);
            est += (double)sig_count(sx, dl_guess) * (double)sig_count(sy, dl_guess) * (X == Y ? 0.5 : 1.0) * (fp + fd);
          }
        keyed.push_back({est, {X, Y}});
      }
    // measured (gpurun r3f): by expected cost for a 1/8 share 4.60 -> 4.45 ms ((H2O)32), 10.80 -> 10.70 ms (C40H82); on whole
    // builds no difference, and on benzene (every kernel latency-bound) 2.6 -> 2.9 ms: used for sharded builds only
    const int order = e.launch_order >= 0 ? e.launch_order : (nranks > 1 ? 2 : 0);
    if (order == 1) std::reverse(keyed.begin(), keyed.end());
    else if (order == 2)
      std::stable_sort(keyed.begin(), keyed.end(), [](const auto &a, const auto &c) { return a.first > c.first; });
    for (const auto &k : keyed) xy_order.push_back(k.second);
  }
  for (const auto &xy : xy_order) {
   const int X = xy.first, Y = xy.second;
    for (int sx = cls_seg[X + 1] - 1; sx >= cls_seg[X]; --sx)
     for (int sy = (X == Y ? sx : cls_seg[Y + 1] - 1); sy >= cls_seg[Y]; --sy) {
      const int qc = qclass_of(X, Y);
      // the kernels see the whole segments; the significant prefixes below only size the grid
      prm.bra_off = b->seg_off[sx]; prm.bra_n = b->seg_off[sx + 1] - b->seg_off[sx];
      prm.ket_off = b->seg_off[sy]; prm.ket_n = b->seg_off[sy + 1] - b->seg_off[sy];
      if (prm.bra_n == 0 || prm.ket_n == 0) continue;
      prm.same = (sx == sy) ? 1 : 0;
      prm.nbx = (prm.ket_n + TILE_KET - 1) / TILE_KET;
      // NEVER skip a launch on the strength of the guess: a class the guess finds empty still gets a minimal grid, whose
      // CTAs either exit on the true bound or (guess too low) walk the rows with the grid stride
      const int bra_sig = std::max(1, sig_count(sx, dl_guess));
      int ket_eff = sig_count(sy, dl_guess);
      if (screen && (int64_t)tlog - dl_guess > -40000)
        ket_eff = std::min(ket_eff, prefix_count(b, sy, tlog - dl_guess - b->h_pair_qls[b->seg_off[sx]]));
      ket_eff = std::max(1, ket_eff);
      if (prm.same) ket_eff = std::min(ket_eff, bra_sig);
      // unscheduled: tile rows are dealt round-robin over all ranks (the rank that receives row 0, the heaviest,
      // rotates with the class pair).  scheduled: this rank owns whole pieces (row subsets p, p+np, ...) of few classes
      std::vector<std::pair<int, int>> pieces;
      if (scheduled) pieces = my_pieces[qc];
      else pieces.push_back({(rank + qc) % nranks, nranks});
     for (const auto &piece : pieces) {
      const int vrank = piece.first, vnranks = piece.second;
      prm.rank = vrank;
      prm.nranks = vnranks;
      const int tile_bra = tile_bra_of(X, Y);
      const int nrows = (bra_sig + tile_bra - 1) / tile_bra;
      const int nrows_local = vrank < nrows ? (nrows - vrank + vnranks - 1) / vnranks : 0;
      if (nrows_local == 0) continue;
      const int nbx_eff = (ket_eff + TILE_KET - 1) / TILE_KET;
      // CTAs per bra tile row (ysplit): CTA y of a row walks ket tiles y, y+ysplit, ...  Enough CTAs for ~256 per SM
      // (rows differ a lot in length), but >= 8 tiles per CTA when that still leaves >= 64 CTAs per SM, so the row's
      // bra data are read once per many tiles.  (QLB_YS_TARGET / QLB_YS_TILES / QLB_YSPLIT_FULL: tuning knobs read in
      // qlb_init; the result was insensitive to them: 65.3-65.9 ms over a 16x range.)
      int ysplit = (int)std::min<int64_t>(nbx_eff, ((int64_t)e.sm_count * ys_target_of(X, Y, e) + nrows_local - 1) / nrows_local);
      const int ys_amort = std::max(1, nbx_eff / e.ys_tiles);
      if ((int64_t)nrows_local * ys_amort >= (int64_t)e.sm_count * 64) ysplit = std::min(ysplit, ys_amort);
      ysplit = std::max(1, ysplit);
      if (e.ysplit_full) ysplit = nbx_eff;
      prm.ysplit = ysplit;
      // staging the bra shells' rows of the density-bound table (2 x nsh entries per warp) pays from a handful of tiles per
      // walk; with short walks (small systems, many ranks) the per-CTA staging cost exceeds the gathers it saves
      prm.dl_smem = (dl_smem_base && nbx_eff / ysplit >= e.dlsmem_min_tiles) ? 1 : 0;
      const int64_t grid = (int64_t)nrows_local * ysplit;
      prm.counters = b->d_counters + 2 * qc;
      if (prof && !qc_started[qc]) cudaEventRecord(b->ev[4 + 2 * qc], s);
      qc_started[qc] = true;
      // deal the class launches over the build stream and the side streams (not while timing classes individually)
      cudaStream_t ls = s;
      if (!prof) {
        if (!forked) {
          QLB_CUDA(cudaEventRecord(e.ev_fork, s));
          for (int i = 0; i < e.nside; ++i) QLB_CUDA(cudaStreamWaitEvent(e.side_stream[i], e.ev_fork, 0));
          forked = true;
        }
        const int which = nlaunch_classes++ % (e.nside + 1);
        if (which > 0) ls = e.side_stream[which - 1];
      }
      const int ce = launch_class_kernel(X, Y, 0, (unsigned)grid, ls, prm);
      if (ce) {
        if (forked)  // join the side streams before reporting the failure
          for (int i = 0; i < e.nside; ++i) {
            cudaEventRecord(e.ev_join[i], e.side_stream[i]);
            cudaStreamWaitEvent(s, e.ev_join[i], 0);
          }
        return cuda_fail((cudaError_t)ce, "eri_class_kernel");
      }
      if (prof) cudaEventRecord(b->ev[4 + 2 * qc + 1], s);
      launches += launch_count_of(X, Y);
     }  // pieces
    }
  }  // class pairs
  if (forked)
    for (int i = 0; i < e.nside; ++i) {
      QLB_CUDA(cudaEventRecord(e.ev_join[i], e.side_stream[i]));
      QLB_CUDA(cudaStreamWaitEvent(s, e.ev_join[i], 0));
    }
  if (stats || calibrate) {
    cudaEventRecord(b->ev[2], s);
    unsigned long long h[2 * NQCLASS];
    QLB_CUDA(cudaMemcpyAsync(h, b->d_counters, sizeof(h), cudaMemcpyDeviceToHost, s));
    QLB_CUDA(cudaStreamSynchronize(s));
    float msq[NQCLASS] = {0};
    for (int qc = 0; qc < NQCLASS; ++qc)
      if (prof && qc_started[qc]) cudaEventElapsedTime(&msq[qc], b->ev[4 + 2 * qc], b->ev[4 + 2 * qc + 1]);
    if (calibrate) {  // this rank's class times ride to the other ranks in the tail of the [J;K] all-reduce
      for (int qc = 0; qc < NQCLASS; ++qc) {
        b->h_ms_local[qc] = msq[qc];
        b->calib_ms[qc] = b->calib_work[qc] = 0.0;
      }
      b->ms_local_pending = true;
      b->calib_nranks = nranks;
    }
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->quartets_unique = b->npairs * (b->npairs + 1) / 2;
      stats->pairs_total = b->npairs;
      const int dlmax_now = screen ? *b->h_dlmax : 0;  // this build's bound (the stream has been synchronised)
      for (int sg = 0; sg < nseg; ++sg) stats->pairs_significant += sig_count(sg, dlmax_now);
      stats->kernel_launches = launches;
      for (int X = 0; X < NCLASS; ++X)
        for (int Y = 0; Y <= X; ++Y) {
          const int qc = qclass_of(X, Y);
          stats->quartets_kept_class[qc] = (int64_t)h[2 * qc];
          stats->prim_quartets_class[qc] = (int64_t)h[2 * qc + 1];
          stats->quartets_kept += (int64_t)h[2 * qc];
          double fp, fd;
          class_flops(X, Y, fp, fd);
          stats->flops_algorithmic += fp * (double)h[2 * qc + 1] + fd * (double)h[2 * qc];
          if (prof && h[2 * qc]) stats->ms_class[qc] = msq[qc];
        }
      float ms = 0;
      cudaEventElapsedTime(&ms, b->ev[0], b->ev[2]);
      stats->ms_total = ms;
      cudaEventElapsedTime(&ms, b->ev[1], b->ev[2]);
      stats->ms_eri = ms;
    }
  }
  return QLB_OK;
}

extern "C" int qlb_symmetrize_device(qlb_basis *b, double *dJ, double *dK, void *stream_) {
  if (int rc = enter(b)) return rc;
  if (!b) return QLB_ERR_ARG;
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : g_engine.stream;
  const int N = b->nbf;
  dim3 block(32, 8), grid((N + 31) / 32, (N + 7) / 8);
  if (dJ) symmetrize_kernel<<<grid, block, 0, s>>>(N, dJ);
  if (dK) symmetrize_kernel<<<grid, block, 0, s>>>(N, dK);
  QLB_CUDA(cudaGetLastError());
  return QLB_OK;
}

extern "C" int qlb_fock_device(qlb_basis *b, const double *dH, const double *dJ, const double *dK, double *dF, void *stream_) {
  if (int rc = enter(b)) return rc;
  if (!b || !dH || !dJ || !dK || !dF) return QLB_ERR_ARG;
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : g_engine.stream;
  const size_t n = (size_t)b->nbf * b->nbf;
  fock_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, dH, dJ, dK, dF);
  QLB_CUDA(cudaGetLastError());
  return QLB_OK;
}

// shared host-buffer path: P up, fused J/K (sharded + all-reduced when a communicator exists), symmetrise
static int jk_host_common(qlb_basis *b, const double *P, double tau, qlb_stats *stats) {
  Engine &e = g_engine;
  const size_t N2 = (size_t)b->nbf * b->nbf;
  QLB_CUDA(cudaSetDevice(e.device));
  QLB_CUDA(cudaMemcpyAsync(b->d_P, P, N2 * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  if (int rc = qlb_build_jk_device(b, b->d_P, tau, b->d_JK, b->d_JK + N2, e.rank, e.nranks, e.stream, stats)) return rc;
  if (e.nranks > 1)
    if (int rc = qlb_allreduce_jk_device(b, b->d_JK, e.stream)) return rc;
  if (int rc = qlb_symmetrize_device(b, b->d_JK, b->d_JK + N2, e.stream)) return rc;
  if (stats) stats->kernel_launches += 2;
  return QLB_OK;
}

extern "C" int qlb_build_jk(qlb_basis *b, const double *P, double tau, double *J, double *K, qlb_stats *stats) {
  if (int rc = enter(b)) return rc;
  if (!b || !P || (!J && !K)) {
    set_error("qlb_build_jk: null pointer");
    return QLB_ERR_ARG;
  }
  Engine &e = g_engine;
  const size_t N2 = (size_t)b->nbf * b->nbf;
  if (int rc = jk_host_common(b, P, tau, stats)) return rc;
  if (J) QLB_CUDA(cudaMemcpyAsync(J, b->d_JK, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  if (K) QLB_CUDA(cudaMemcpyAsync(K, b->d_JK + N2, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  return QLB_OK;
}

extern "C" int qlb_fock(qlb_basis *b, const double *H, const double *P, double tau, double *F, qlb_stats *stats) {
  if (int rc = enter(b)) return rc;
  if (!b || !H || !P || !F) {
    set_error("qlb_fock: null pointer");
    return QLB_ERR_ARG;
  }
  Engine &e = g_engine;
  const size_t N2 = (size_t)b->nbf * b->nbf;
  QLB_CUDA(cudaSetDevice(e.device));
  // H is only needed when F is assembled: its upload runs beside the build on the copy stream (the previous call has
  // been synchronised, so d_H is free)
  QLB_CUDA(cudaMemcpyAsync(b->d_H, H, N2 * sizeof(double), cudaMemcpyHostToDevice, e.copy_stream));
  QLB_CUDA(cudaEventRecord(e.ev_copy, e.copy_stream));
  if (int rc = jk_host_common(b, P, tau, stats)) return rc;
  QLB_CUDA(cudaStreamWaitEvent(e.stream, e.ev_copy, 0));
  // F overwrites the P scratch (P is no longer needed)
  if (int rc = qlb_fock_device(b, b->d_H, b->d_JK, b->d_JK + N2, b->d_P, e.stream)) return rc;
  if (stats) stats->kernel_launches += 1;
  QLB_CUDA(cudaMemcpyAsync(F, b->d_P, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  QLB_CUDA(cudaStreamSynchronize(e.stream));
  return QLB_OK;
}

// =================================================================== ERI tensor / block
static int tensor_device(qlb_basis *b, double *dT) {
  Engine &e = g_engine;
  const size_t N = b->nbf;
  QLB_CUDA(cudaMemsetAsync(dT, 0, N * N * N * N * sizeof(double), e.stream));
  ClassParams prm;
  base_params(b, prm);
  prm.tensor = dT;
  for (int X = 0; X < NCLASS; ++X)
    for (int Y = 0; Y <= X; ++Y) {
      prm.bra_off = b->class_off[X]; prm.bra_n = b->class_off[X + 1] - b->class_off[X];
      prm.ket_off = b->class_off[Y]; prm.ket_n = b->class_off[Y + 1] - b->class_off[Y];
      if (prm.bra_n == 0 || prm.ket_n == 0) continue;
      prm.same = (X == Y) ? 1 : 0;
      prm.nbx = (prm.ket_n + TILE_KET - 1) / TILE_KET;
      prm.ysplit = prm.nbx;
      const int tile_bra = tile_bra_of(X, Y);
      const int64_t grid = (int64_t)((prm.bra_n + tile_bra - 1) / tile_bra) * prm.nbx;
      if (int ce = launch_class_kernel(X, Y, 1, (unsigned)grid, e.stream, prm)) return cuda_fail((cudaError_t)ce, "eri_class_kernel(tensor)");
    }
  return QLB_OK;
}

extern "C" int qlb_eri_tensor(qlb_basis *b, double *out) {
  if (int rc = enter(b)) return rc;
  if (!b || !out) return QLB_ERR_ARG;
  if (b->nbf > 160) {
    set_error("qlb_eri_tensor: N > 160; the N^4 tensor is only materialised for test-sized systems");
    return QLB_ERR_ARG;
  }
  Engine &e = g_engine;
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N = b->nbf, n4 = N * N * N * N;
  double *dT = nullptr;
  QLB_CUDA(cudaMalloc((void **)&dT, n4 * sizeof(double)));
  int rc = tensor_device(b, dT);
  if (rc == QLB_OK) {
    cudaError_t ce = cudaMemcpyAsync(out, dT, n4 * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "tensor D2H");
  }
  cudaFree(dT);
  return rc;
}

extern "C" int qlb_eri_block(qlb_basis *b, int i, int j, int k, int l, double *out) {
  if (int rc = enter(b)) return rc;
  if (!b || !out) return QLB_ERR_ARG;
  const int idx[4] = {i, j, k, l};
  for (int s = 0; s < 4; ++s)
    if (idx[s] < 0 || idx[s] >= b->nsh) {
      set_error("qlb_eri_block: shell index out of range");
      return QLB_ERR_ARG;
    }
  // a four-shell sub-basis; its full tensor contains the requested block
  std::vector<double> xyz, ex, co;
  std::vector<int32_t> L, np;
  for (int s = 0; s < 4; ++s) {
    const int sh = idx[s];
    xyz.insert(xyz.end(), b->h_xyz.begin() + 3 * sh, b->h_xyz.begin() + 3 * sh + 3);
    L.push_back(b->h_L[sh]);
    np.push_back(b->h_nprim[sh]);
    ex.insert(ex.end(), b->h_exps.begin() + b->h_poff[sh], b->h_exps.begin() + b->h_poff[sh + 1]);
    co.insert(co.end(), b->h_coefs.begin() + b->h_poff[sh], b->h_coefs.begin() + b->h_poff[sh + 1]);
  }
  qlb_basis *mini = nullptr;
  if (int rc = qlb_basis_create(4, xyz.data(), L.data(), np.data(), ex.data(), co.data(), &mini)) return rc;
  const size_t n = mini->nbf;
  std::vector<double> T(n * n * n * n);
  int rc = qlb_eri_tensor(mini, T.data());
  if (rc == QLB_OK) {
    const int nA = ncart(L[0]), nB = ncart(L[1]), nC = ncart(L[2]), nD = ncart(L[3]);
    const int *bo = mini->h_boff.data();
    for (int d = 0; d < nD; ++d)
      for (int c = 0; c < nC; ++c)
        for (int bb = 0; bb < nB; ++bb)
          for (int a = 0; a < nA; ++a)
            out[a + nA * (bb + nB * (c + nC * d))] = T[(bo[0] + a) + n * ((bo[1] + bb) + n * ((bo[2] + c) + n * (size_t)(bo[3] + d)))];
  }
  qlb_basis_destroy(mini);
  return rc;
}

// =================================================================== NCCL (loaded lazily; one process per GPU)
namespace {
typedef ncclResult_t (*fn_GetUniqueId)(ncclUniqueId *);
typedef ncclResult_t (*fn_CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_CommDestroy)(ncclComm_t);
typedef const char *(*fn_GetErrorString)(ncclResult_t);
typedef ncclResult_t (*fn_CommInitAll)(ncclComm_t *, int, const int *);
typedef ncclResult_t (*fn_Group)(void);
typedef ncclResult_t (*fn_Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
void *g_nccl_lib = nullptr;
fn_GetUniqueId p_GetUniqueId;
fn_CommInitRank p_CommInitRank;
fn_AllReduce p_AllReduce;
fn_CommDestroy p_CommDestroy;
fn_GetErrorString p_GetErrorString;
fn_CommInitAll p_CommInitAll;
fn_Group p_GroupStart, p_GroupEnd;
fn_Broadcast p_Broadcast;

int load_nccl() {
  Engine &e = g_engine;
  if (p_AllReduce) {  // already resolved (by another device's engine)
    if (!e.nccl_lib) e.nccl_lib = g_nccl_lib;
    return QLB_OK;
  }
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    e.nccl_lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (e.nccl_lib) break;
  }
  if (!e.nccl_lib) {
    set_error(std::string("cannot dlopen libnccl: ") + dlerror());
    return QLB_ERR_NCCL;
  }
  p_GetUniqueId = (fn_GetUniqueId)dlsym(e.nccl_lib, "ncclGetUniqueId");
  p_CommInitRank = (fn_CommInitRank)dlsym(e.nccl_lib, "ncclCommInitRank");
  p_AllReduce = (fn_AllReduce)dlsym(e.nccl_lib, "ncclAllReduce");
  p_CommDestroy = (fn_CommDestroy)dlsym(e.nccl_lib, "ncclCommDestroy");
  p_GetErrorString = (fn_GetErrorString)dlsym(e.nccl_lib, "ncclGetErrorString");
  p_CommInitAll = (fn_CommInitAll)dlsym(e.nccl_lib, "ncclCommInitAll");
  p_GroupStart = (fn_Group)dlsym(e.nccl_lib, "ncclGroupStart");
  p_GroupEnd = (fn_Group)dlsym(e.nccl_lib, "ncclGroupEnd");
  p_Broadcast = (fn_Broadcast)dlsym(e.nccl_lib, "ncclBroadcast");
  if (!p_GetUniqueId || !p_CommInitRank || !p_AllReduce || !p_CommDestroy || !p_CommInitAll || !p_GroupStart || !p_GroupEnd || !p_Broadcast) {
    set_error("libnccl lacks a required symbol");
    return QLB_ERR_NCCL;
  }
  g_nccl_lib = e.nccl_lib;
  return QLB_OK;
}
int nccl_fail(ncclResult_t r, const char *what) {
  set_error(std::string("NCCL error in ") + what + ": " + (p_GetErrorString ? p_GetErrorString(r) : "?"));
  return QLB_ERR_NCCL;
}
}  // namespace

extern "C" int qlb_comm_unique_id(void *id128) {
  if (!id128) return QLB_ERR_ARG;
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  ncclResult_t r = p_GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return QLB_OK;
}

extern "C" int qlb_comm_init(const void *id128, int rank, int nranks) {
  if (int rc = require_ready()) return rc;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return QLB_ERR_ARG;
  if (int rc = load_nccl()) return rc;
  Engine &e = g_engine;
  QLB_CUDA(cudaSetDevice(e.device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  ncclResult_t r = p_CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  e.nccl_comm = (void *)comm;
  e.rank = rank;
  e.nranks = nranks;
  return QLB_OK;
}

extern "C" int qlb_comm_destroy(void) {
  Engine &e = g_engine;
  if (e.nccl_comm && p_CommDestroy) p_CommDestroy((ncclComm_t)e.nccl_comm);
  e.nccl_comm = nullptr;
  e.rank = 0;
  e.nranks = 1;
  return QLB_OK;
}

// fills the QLB_JK_TAIL doubles behind [J;K]: per-class counters of this build and, after a calibration build, this
// rank's class kernel times
__global__ void jk_tail_kernel(const unsigned long long *__restrict__ c, const double *__restrict__ ms, double *tail) {
  const int i = threadIdx.x;
  if (i < 2 * NQCLASS) tail[i] = (double)c[i];
  else if (i < 3 * NQCLASS) tail[i] = ms ? ms[i - 2 * NQCLASS] : 0.0;
  else if (i < QLB_JK_TAIL) tail[i] = 0.0;
}

int qlb::allreduce_sum_device(double *d, size_t n, cudaStream_t s) {
  Engine &e = g_engine;
  if (!e.nccl_comm) {
    set_error("allreduce: no communicator (qlb_comm_init)");
    return QLB_ERR_STATE;
  }
  ncclResult_t r = p_AllReduce(d, d, n, ncclDouble, ncclSum, (ncclComm_t)e.nccl_comm, s);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  return QLB_OK;
}

// The all-reduce of [J;K] + schedule feedback in three stream-ordered parts, so that the single-process front end can
// put the collectives of all its devices into one NCCL group (inside a group the collective is only enqueued at
// ncclGroupEnd: anything that must follow it on the stream has to be issued after the group).
static int allreduce_jk_pre(qlb_basis *b, double *dJK, cudaStream_t s) {
  static_assert(3 * NQCLASS <= QLB_JK_TAIL, "QLB_JK_TAIL too small");
  const size_t n = 2 * (size_t)b->nbf * b->nbf;
  // schedule feedback rides behind [J;K]: ONE collective per build
  const double *dms = nullptr;
  if (b->ms_local_pending) {
    QLB_CUDA(cudaMemcpyAsync(b->d_ms_local, b->h_ms_local, NQCLASS * sizeof(double), cudaMemcpyHostToDevice, s));
    dms = b->d_ms_local;
    b->ms_local_pending = false;
  }
  jk_tail_kernel<<<1, QLB_JK_TAIL, 0, s>>>(b->d_counters, dms, dJK + n);
  QLB_CUDA(cudaGetLastError());
  return QLB_OK;
}
static int allreduce_jk_reduce(qlb_basis *b, double *dJK, cudaStream_t s) {
  const size_t n = 2 * (size_t)b->nbf * b->nbf;
  ncclResult_t r = p_AllReduce(dJK, dJK, n + QLB_JK_TAIL, ncclDouble, ncclSum, (ncclComm_t)g_engine.nccl_comm, s);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  return QLB_OK;
}
static int allreduce_jk_post(qlb_basis *b, double *dJK, cudaStream_t s) {
  const size_t n = 2 * (size_t)b->nbf * b->nbf;
  QLB_CUDA(cudaMemcpyAsync(b->h_gcounts, dJK + n, 3 * NQCLASS * sizeof(double), cudaMemcpyDeviceToHost, s));
  QLB_CUDA(cudaEventRecord(b->ev_gcounts, s));
  b->gcounts_pending = true;
  b->gcounts_nranks = g_engine.nranks;
  return QLB_OK;
}

extern "C" int qlb_allreduce_jk_device(qlb_basis *b, double *dJK, void *stream_) {
  if (int rc = enter(b)) return rc;
  Engine &e = g_engine;
  if (!b || !dJK) return QLB_ERR_ARG;
  if (!e.nccl_comm) {
    set_error("qlb_allreduce_jk_device: no communicator (qlb_comm_init)");
    return QLB_ERR_STATE;
  }
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : e.stream;
  if (int rc = allreduce_jk_pre(b, dJK, s)) return rc;
  if (int rc = allreduce_jk_reduce(b, dJK, s)) return rc;
  return allreduce_jk_post(b, dJK, s);
}

// =================================================================== single process, several devices
// One host thread drives ngpu devices: a qlb_basis per device, a communicator per device from ncclCommInitAll, every
// device evaluates its share of the quartets of the SAME build, one grouped all-reduce of [J;K] combines them.  This is
// the mode a single Julia process (the reference's evaluateSCF caller) uses to get the whole box.
struct qlb_multi {
  int ngpu = 0;
  std::vector<qlb_basis *> basis;
  std::vector<ncclComm_t> comms;
};

extern "C" int qlb_multi_destroy(qlb_multi *m) {
  if (!m) return QLB_OK;
  for (int d = 0; d < m->ngpu; ++d) {
    if (d < (int)m->basis.size() && m->basis[d]) qlb_basis_destroy(m->basis[d]);
    if (d < (int)m->comms.size() && m->comms[d]) {
      if (p_CommDestroy) p_CommDestroy(m->comms[d]);
      g_engines[d].nccl_comm = nullptr;
      g_engines[d].rank = 0;
      g_engines[d].nranks = 1;
    }
  }
  delete m;
  if (g_engines[0].ready) use_engine(&g_engines[0]);
  return QLB_OK;
}

extern "C" int qlb_multi_create(int ngpu, int nsh, const double *centers, const int32_t *L, const int32_t *nprim, const double *exps,
                                const double *coefs, qlb_multi **out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available: libqlb200 has no CPU fallback");
    return QLB_ERR_CUDA;
  }
  if (!out || ngpu < 1 || ngpu > ndev || ngpu > MAX_DEVICES) {
    set_error("qlb_multi_create: ngpu must be between 1 and the number of visible devices");
    return QLB_ERR_ARG;
  }
  qlb_multi *m = new qlb_multi();
  m->ngpu = ngpu;
  m->basis.assign(ngpu, nullptr);
  m->comms.assign(ngpu, nullptr);
  for (int d = 0; d < ngpu; ++d) {
    int rc = qlb_init(d);
    if (!rc) rc = qlb_basis_create(nsh, centers, L, nprim, exps, coefs, &m->basis[d]);
    if (rc) {
      qlb_multi_destroy(m);
      return rc;
    }
  }
  if (ngpu > 1) {
    use_engine(&g_engines[0]);
    if (int rc = load_nccl()) {
      qlb_multi_destroy(m);
      return rc;
    }
    std::vector<int> devs(ngpu);
    for (int d = 0; d < ngpu; ++d) devs[d] = d;
    ncclResult_t r = p_CommInitAll(m->comms.data(), ngpu, devs.data());
    if (r != ncclSuccess) {
      qlb_multi_destroy(m);
      return nccl_fail(r, "ncclCommInitAll");
    }
    for (int d = 0; d < ngpu; ++d) {
      g_engines[d].nccl_comm = (void *)m->comms[d];
      g_engines[d].rank = d;
      g_engines[d].nranks = ngpu;
    }
  }
  *out = m;
  return QLB_OK;
}

extern "C" int qlb_multi_ngpu(const qlb_multi *m, int *ngpu) {
  if (!m || !ngpu) return QLB_ERR_ARG;
  *ngpu = m->ngpu;
  return QLB_OK;
}

// P (and H) go up to device 0 once and travel to the other devices over NVLink (grouped ncclBroadcast); every device
// enqueues its share of the build; one grouped all-reduce; device 0 symmetrises / assembles F and answers the host.
static int multi_build(qlb_multi *m, const double *H, const double *P, double tau, double *J, double *K, double *F, qlb_stats *stats) {
  const int n = m->ngpu;
  qlb_basis *b0 = m->basis[0];
  const size_t N2 = (size_t)b0->nbf * b0->nbf;
  use_engine(&g_engines[0]);
  QLB_CUDA(cudaMemcpyAsync(b0->d_P, P, N2 * sizeof(double), cudaMemcpyHostToDevice, g_engines[0].stream));
  if (H) QLB_CUDA(cudaMemcpyAsync(b0->d_H, H, N2 * sizeof(double), cudaMemcpyHostToDevice, g_engines[0].stream));
  if (n > 1) {
    p_GroupStart();
    for (int d = 0; d < n; ++d) {
      use_engine(&g_engines[d]);
      ncclResult_t r = p_Broadcast(b0->d_P, m->basis[d]->d_P, N2, ncclDouble, 0, m->comms[d], g_engines[d].stream);
      if (r != ncclSuccess) {
        p_GroupEnd();
        return nccl_fail(r, "ncclBroadcast(P)");
      }
    }
    ncclResult_t r = p_GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd(broadcast)");
  }
  // the other devices first: device 0's call may wait for its stream (statistics)
  for (int k = 0; k < n; ++k) {
    const int d = (k + 1) % n;
    qlb_basis *b = m->basis[d];
    if (int rc = qlb_build_jk_device(b, b->d_P, tau, b->d_JK, b->d_JK + N2, d, n, nullptr, d == 0 ? stats : nullptr)) return rc;
  }
  if (n > 1) {
    for (int d = 0; d < n; ++d) {
      use_engine(&g_engines[d]);
      if (int rc = allreduce_jk_pre(m->basis[d], m->basis[d]->d_JK, g_engines[d].stream)) return rc;
    }
    p_GroupStart();
    for (int d = 0; d < n; ++d) {
      use_engine(&g_engines[d]);
      if (int rc = allreduce_jk_reduce(m->basis[d], m->basis[d]->d_JK, g_engines[d].stream)) {
        p_GroupEnd();
        return rc;
      }
    }
    ncclResult_t r = p_GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd(all-reduce)");
    for (int d = 0; d < n; ++d) {  // after the group: the collective is on the streams now
      use_engine(&g_engines[d]);
      if (int rc = allreduce_jk_post(m->basis[d], m->basis[d]->d_JK, g_engines[d].stream)) return rc;
    }
  }
  use_engine(&g_engines[0]);
  Engine &e = g_engines[0];
  if (int rc = qlb_symmetrize_device(b0, b0->d_JK, b0->d_JK + N2, e.stream)) return rc;
  if (F) {
    if (int rc = qlb_fock_device(b0, b0->d_H, b0->d_JK, b0->d_JK + N2, b0->d_P, e.stream)) return rc;  // F overwrites the P scratch
    QLB_CUDA(cudaMemcpyAsync(F, b0->d_P, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  }
  if (J) QLB_CUDA(cudaMemcpyAsync(J, b0->d_JK, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  if (K) QLB_CUDA(cudaMemcpyAsync(K, b0->d_JK + N2, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  for (int d = n - 1; d >= 0; --d) {
    use_engine(&g_engines[d]);
    QLB_CUDA(cudaStreamSynchronize(g_engines[d].stream));
  }
  return QLB_OK;
}

extern "C" int qlb_multi_build_jk(qlb_multi *m, const double *P, double tau, double *J, double *K, qlb_stats *stats) {
  if (!m || !P || (!J && !K)) {
    set_error("qlb_multi_build_jk: null pointer");
    return QLB_ERR_ARG;
  }
  return multi_build(m, nullptr, P, tau, J, K, nullptr, stats);
}

extern "C" int qlb_multi_fock(qlb_multi *m, const double *H, const double *P, double tau, double *F, qlb_stats *stats) {
  if (!m || !H || !P || !F) {
    set_error("qlb_multi_fock: null pointer");
    return QLB_ERR_ARG;
  }
  return multi_build(m, H, P, tau, nullptr, nullptr, F, stats);
}

// the device-0 handle of a multi-device object: one-electron matrices, Schwarz bounds, ... come from it
extern "C" int qlb_multi_basis0(qlb_multi *m, qlb_basis **b) {
  if (!m || !b) return QLB_ERR_ARG;
  *b = m->basis[0];
  return QLB_OK;
}

// =================================================================== one-electron matrices
static int one_electron_host(qlb_basis *b, int which, int natoms, const double *xyz, const double *Z, double *out) {
  if (int rc = enter(b)) return rc;
  if (!b || !out || (which == 2 && (natoms <= 0 || !xyz || !Z))) return QLB_ERR_ARG;
  Engine &e = g_engine;
  QLB_CUDA(cudaSetDevice(e.device));
  const size_t N2 = (size_t)b->nbf * b->nbf;
  double *d_at = nullptr;
  if (which == 2) {
    std::vector<double> at(4 * (size_t)natoms);
    for (int i = 0; i < natoms; ++i) {
      at[4 * i] = xyz[3 * i]; at[4 * i + 1] = xyz[3 * i + 1]; at[4 * i + 2] = xyz[3 * i + 2]; at[4 * i + 3] = Z[i];
    }
    QLB_CUDA(cudaMalloc((void **)&d_at, at.size() * sizeof(double)));
    QLB_CUDA(cudaMemcpy(d_at, at.data(), at.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  int rc = one_electron_device(b, which, natoms, d_at, b->d_H, e.stream);
  if (rc == QLB_OK) {
    cudaError_t ce = cudaMemcpyAsync(out, b->d_H, N2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e.stream);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "one-electron D2H");
  }
  if (d_at) cudaFree(d_at);
  return rc;
}
extern "C" int qlb_overlap(qlb_basis *b, double *S) { return one_electron_host(b, 0, 0, nullptr, nullptr, S); }
extern "C" int qlb_kinetic(qlb_basis *b, double *T) { return one_electron_host(b, 1, 0, nullptr, nullptr, T); }
extern "C" int qlb_nuclear_attraction(qlb_basis *b, int natoms, const double *xyz, const double *Z, double *V) {
  return one_electron_host(b, 2, natoms, xyz, Z, V);
}
