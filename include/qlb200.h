/*
 * qlb200.h -- C ABI of libqlb200.so, the B200-native Fock-build engine behind QuantumLab.jl's SCF hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and returns an int status (0 = ok,
 * non-zero = error; qlb_last_error() gives the message).  There is NO CPU fallback: every compute entry
 * point fails with QLB_ERR_CUDA if no sm_100-class device is usable.
 *
 * The reference interface each entry point replaces is QuantumLab.jl's only FFI seam, the libint2 shim
 * (deps/usr/src/libint2jl.cpp, bound by src/SubModules/LibInt2Module.jl), plus the Julia functions that
 * sit directly above it on the hot path.  Citations are file:line relative to the reference checkout.
 *
 * Conventions (identical to the reference's native Julia path):
 *   - all matrices / tensors are column-major double, first index fastest, in the reference's basis
 *     function order: shells in the order given, Cartesian components x-descending then y-descending
 *     (BaseModule.jl:110-122), e.g. d = xx,xy,xz,yy,yz,zz;
 *   - shell coefficients arrive already normalised the way the Shell constructor leaves them
 *     (ShellModule.jl:25-48); the non-axial factors of ShellModule.jl:50-58 (d_xy: sqrt 3) are applied
 *     on the device, so outputs equal the native Julia values directly;
 *   - the density P is the reference's P = Cocc*Cocc' (no factor 2, HartreeFockModule.jl:82), symmetric.
 */
#ifndef QLB200_H
#define QLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLB_OK 0
#define QLB_ERR_ARG 1       /* bad argument (null pointer, L > QLB_MAX_L, ...)            */
#define QLB_ERR_CUDA 2      /* CUDA runtime / launch failure, or no usable device           */
#define QLB_ERR_STATE 3     /* qlb_init not called, handle destroyed, ...                   */
#define QLB_ERR_NCCL 4      /* collective failure                                           */
#define QLB_ERR_NOMEM 5

#define QLB_MAX_L 2         /* highest ORBITAL shell angular momentum with ERI class kernels (s,p,d) */
#define QLB_MAX_L_AUX 3     /* highest AUXILIARY (RI) shell angular momentum (s,p,d,f)           */
#define QLB_NCLASS 6        /* shell-pair classes ss,ps,pp,ds,dp,dd                          */
#define QLB_NQCLASS 21      /* canonical (bra class >= ket class) quartet classes            */
#define QLB_JK_TAIL 64      /* doubles behind [J;K] that carry the schedule feedback through the one all-reduce */

typedef struct qlb_basis qlb_basis; /* opaque: device-resident shells, shell pairs, Schwarz bounds */

/* Per-build statistics (all counts are exact integers; times are CUDA-event milliseconds). */
typedef struct qlb_stats {
  int64_t quartets_unique;            /* unique shell quartets npair*(npair+1)/2 (before screening)        */
  int64_t quartets_kept;              /* quartets that passed the integer Schwarz x density rule            */
  int64_t quartets_kept_class[QLB_NQCLASS];      /* the same, per canonical L-class                         */
  int64_t prim_quartets_class[QLB_NQCLASS];      /* primitive quartets evaluated, per class                 */
  int64_t pairs_total;                /* unique shell pairs                                                  */
  int64_t pairs_significant;          /* pairs that can survive with the current tau and density            */
  double ms_total;                    /* whole call, on the device timeline (events on the build stream)     */
  double ms_eri;                      /* the ERI/digestion class kernels only                                */
  double ms_class[QLB_NQCLASS];       /* per class kernel (only filled when profiling is switched on)        */
  double flops_algorithmic;           /* SURVEY.md 8(d) minimum-MD flop count for what was evaluated          */
  int32_t kernel_launches;            /* kernels launched by this call                                       */
  int32_t reserved;
} qlb_stats;

/* ---- lifetime -----------------------------------------------------------------------------------------
 * replaces libint2start()/libint2stop()  (libint2jl.cpp:3-9, LibInt2Module.jl:30-40).
 * device < 0 selects the current device.  One engine per device; handles remember their engine, handle-less calls use the
 * engine of the last qlb_init.  One process per GPU (torchrun) and one process driving several GPUs (qlb_multi_*) both work. */
int qlb_init(int device);
int qlb_finalize(void);
const char *qlb_last_error(void);
int qlb_device_info(char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor);

/* ---- basis ---------------------------------------------------------------------------------------------
 * replaces createShell() per shell (libint2jl.cpp:11-24, LibInt2Module.jl:70-104) and
 * createEngineCoulomb()/destroyEngine() (libint2jl.cpp:53-60,88-93): the handle owns the device copy of the
 * shells AND the geometry-only precomputation the reference redoes per primitive quartet
 * (GaussProductFundamental, IntegralsModule.jl:66-83): primitive-pair records, Schwarz bounds, class bins.
 *   centers: 3*nsh (bohr)   L: nsh   nprim: nsh   exps/coefs: concatenated per shell. */
int qlb_basis_create(int nsh, const double *centers, const int32_t *L, const int32_t *nprim,
                     const double *exps, const double *coefs, qlb_basis **out);
int qlb_basis_destroy(qlb_basis *b);
int qlb_basis_nbf(const qlb_basis *b, int *nbf);
int qlb_basis_npairs(const qlb_basis *b, int64_t *npairs);

/* ---- ERI entry points of IntegralsModule -----------------------------------------------------------------
 * qlb_eri_block  replaces compute4cInts() (libint2jl.cpp:100-103) as used by
 *   computeTensorBlockElectronRepulsionIntegrals (LibInt2Module.jl:319-334, SpecialMatricesModule.jl:184-186):
 *   out[mu + nA*(nu + nB*(la + nC*si))] = (mu nu|la si) for shells (i,j,k,l) (0-based), non-axial factors applied.
 * qlb_eri_tensor replaces computeTensorElectronRepulsionIntegrals (SpecialMatricesModule.jl:180-207):
 *   the full N^4 tensor, only for N <= 256 (the engine never needs it; the reference's tests do). */
int qlb_eri_block(qlb_basis *b, int i, int j, int k, int l, double *out);
int qlb_eri_tensor(qlb_basis *b, double *out);

/* Schwarz bounds Q_ij = sqrt(max |(mu nu|mu nu)|) as an nsh x nsh symmetric matrix (new functionality:
 * the reference does not screen). */
int qlb_schwarz(qlb_basis *b, double *Q);

/* ---- J / K / Fock ----------------------------------------------------------------------------------------
 * qlb_build_jk replaces contractTensorBlocksWithMatrix_CoulombLike / _ExchangeLike
 *   (SpecialMatricesModule.jl:33-100) as called by computeMatrixCoulomb / computeMatrixExchange
 *   (:231-244, :268-281): J[mu,nu] = sum (mu nu|la si) P[la,si], K[mu,nu] = sum (mu la|nu si) P[la,si], in ONE
 *   fused pass over the unique, screened quartets.  Host buffers N x N; J or K may be NULL.
 *   A quartet (ab|cd) is evaluated iff qlog(Q_ab) + qlog(Q_cd) + max qlog(|P| over the six shell blocks
 *   ab,cd,ac,ad,bc,bd) >= qlog(tau), qlog(x) = ceil(8*log2 x): integer arithmetic, so counts are reproducible.
 *   tau <= 0 switches screening off.  P MUST be symmetric: the kernels digest one triangle of P and symmetrise J and K
 *   afterwards (the reference's nsh^4 loop would also accept a non-symmetric matrix; the host mirror rejects one).
 * qlb_fock replaces computeMatrixFock (SpecialMatricesModule.jl:284-290): F = H + 2J - K with H = T + V. */
int qlb_build_jk(qlb_basis *b, const double *P, double tau, double *J, double *K, qlb_stats *stats);
int qlb_fock(qlb_basis *b, const double *H, const double *P, double tau, double *F, qlb_stats *stats);

/* Tensor flavour of computeMatrixCoulomb / computeMatrixExchange (SpecialMatricesModule.jl:227-229, 264-266): contracts a
 * STORED N^4 tensor T[mu,nu,la,si] (host, N <= 160) with P on the device.  exchange == 0: J[mu,nu] = sum T[mu,nu,la,si] P[la,si];
 * exchange != 0: K[mu,nu] = sum T[mu,la,nu,si] P[la,si]. */
int qlb_tensor_contract(int N, const double *T, const double *P, int exchange, double *out);

/* Device-resident variant for callers that keep P/J/K in HBM (one process per GPU):
 * dP, dJ, dK are DEVICE pointers (N x N doubles).  Only the quartet tiles owned by `rank` of `nranks` are
 * evaluated, dJ/dK receive the rank's PARTIAL, un-symmetrised sums (J, K as defined above are obtained by
 * summing the partials over ranks and calling qlb_symmetrize_device).  Work is enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = the engine's own stream) and the call returns after enqueueing
 * unless stats != NULL (stats need a synchronisation). */
int qlb_build_jk_device(qlb_basis *b, const double *dP, double tau, double *dJ, double *dK, int rank, int nranks,
                        void *stream, qlb_stats *stats);
int qlb_symmetrize_device(qlb_basis *b, double *dJ, double *dK, void *stream);
/* F = H + 2J - K elementwise on device pointers (N x N) */
int qlb_fock_device(qlb_basis *b, const double *dH, const double *dJ, const double *dK, double *dF, void *stream);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ----------------------------------------------------
 * qlb_comm_unique_id fills a 128-byte NCCL id on rank 0; the host broadcasts it by any means and every rank
 * calls qlb_comm_init.  After that qlb_build_jk / qlb_fock / qlb_build_jk_device(+qlb_allreduce_jk_device)
 * shard the quartet tiles over the ranks and combine the partial [J;K] with ONE ncclAllReduce per build. */
int qlb_comm_unique_id(void *id128);
int qlb_comm_init(const void *id128, int rank, int nranks);
int qlb_comm_destroy(void);
int qlb_allreduce_jk_device(qlb_basis *b, double *dJK /* 2*N*N + QLB_JK_TAIL doubles, contiguous */, void *stream);
/* The first sharded build of a basis deals bra tile rows round-robin and times every class kernel (calibration); later
 * builds deal whole pieces of classes by measured cost: class time of the calibration build, scaled by the class's
 * current work, all-rank sums of which ride in the QLB_JK_TAIL doubles behind [J;K] through the SAME all-reduce.
 * qlb_schedule_pieces exposes that schedule: cost[qc] (any unit) of class pair qc; npieces[qc] pieces (row subsets
 * p, p+np, ..); owner[qc*nranks + p] = rank that evaluates piece p.  Pure host function (tests). */
int qlb_schedule_pieces(const double *cost, int nranks, int32_t *npieces, int32_t *owner);

/* ---- single process, several GPUs: what a single Julia process (the evaluateSCF caller) uses ------------
 * qlb_multi_create     qlb_init + qlb_basis_create on devices 0..ngpu-1 and one communicator per device (ncclCommInitAll).
 * qlb_multi_build_jk / qlb_multi_fock: host buffers as qlb_build_jk / qlb_fock.  P (and H) go up to device 0 once and reach
 *                      the other devices over NVLink (grouped ncclBroadcast); every device evaluates its share of the quartet
 *                      rows of the build; ONE grouped all-reduce of [J;K]; device 0 symmetrises / assembles F and answers.
 * qlb_multi_basis0     the basis handle of device 0, e.g. for qlb_overlap, qlb_schwarz, qlb_scf_create. */
typedef struct qlb_multi qlb_multi;
int qlb_multi_create(int ngpu, int nsh, const double *centers, const int32_t *L, const int32_t *nprim, const double *exps,
                     const double *coefs, qlb_multi **out);
int qlb_multi_destroy(qlb_multi *m);
int qlb_multi_ngpu(const qlb_multi *m, int *ngpu);
int qlb_multi_basis0(qlb_multi *m, qlb_basis **b);
int qlb_multi_build_jk(qlb_multi *m, const double *P, double tau, double *J, double *K, qlb_stats *stats);
int qlb_multi_fock(qlb_multi *m, const double *H, const double *P, double tau, double *F, qlb_stats *stats);

/* ---- one-electron matrices (SpecialMatricesModule.jl:102-177) ---------------------------- */
int qlb_overlap(qlb_basis *b, double *S);
int qlb_kinetic(qlb_basis *b, double *T);
int qlb_nuclear_attraction(qlb_basis *b, int natoms, const double *xyz, const double *Z, double *V);

/* ---- resolution of the identity, Coulomb metric (ResolutionIdentityModule.jl:41-78) -------------------------
 * qlb_ri_create        builds B = (mu nu|P)(P|Q)^-1/2 on the device for an auxiliary shell list (same array conventions as
 *                      qlb_basis_create, shells up to f): replaces computeMatrixResolutionOfTheIdentityCoulombMetric
 *                      (:41-50).  B is stored once, packed over mu >= nu; with a communicator (qlb_comm_init) every rank
 *                      keeps a contiguous slice of the auxiliary index (qlb_ri_slice).
 * qlb_ri_b_tensor      copies B out as the reference's N x N x Naux array (this rank's slice of Q).
 * qlb_ri_build_jk      J[mu,nu] = sum_Q B[mu,nu,Q] sum B[la,si,Q] P[la,si]  (RI-J; new) and
 *                      K[mu,la] = sum_Q sum B[mu,nu,Q] P[nu,si] B[la,si,Q]  = computeMatrixExchangeRIK (:69-78) for ANY
 *                      symmetric P, with the library's own FP64 tensor-core (DMMA) GEMM and streaming GEMV kernels; the
 *                      N^4 tensor of the reference is never formed.  The ranks' partial sums over their Q slices are
 *                      combined by one all-reduce of [J;K].  ms (optional): device time of the contractions.
 * qlb_ri_build_jk_occ  the same for P = Cocc Cocc^T given the occupied orbitals (N x nocc): K costs 3 N^2 nocc Naux flops. */
typedef struct qlb_ri qlb_ri;
int qlb_ri_create(qlb_basis *orbital, int naux_sh, const double *centers, const int32_t *L, const int32_t *nprim,
                  const double *exps, const double *coefs, qlb_ri **out);
int qlb_ri_destroy(qlb_ri *r);
int qlb_ri_naux(const qlb_ri *r, int *naux);
int qlb_ri_slice(const qlb_ri *r, int *q0, int *nq);
int qlb_ri_b_tensor(qlb_ri *r, double *B);
int qlb_ri_eri_tensor(qlb_ri *r, double *out); /* B B^T as N^4 (computeTensorElectronRepulsionIntegralsRICoulomb :54-58), N <= 64 */
int qlb_ri_build_jk(qlb_ri *r, const double *P, double *J, double *K, double *ms);
int qlb_ri_build_jk_occ(qlb_ri *r, const double *Cocc, int nocc, double *J, double *K, double *ms);

/* ---- SCF step on the device (HartreeFockModule.jl:73-84, :14-41; SpecialMatricesModule.jl:284-290) --------
 * A qlb_scf keeps S, T, V, P, J, K, F and the eigenvectors in HBM, so one Roothaan iteration moves four doubles to the
 * host instead of three N x N matrices.
 * qlb_scf_create       uploads S, T, V (N x N, host) once.
 * qlb_scf_set_density  uploads the initial guess (computeDensityMatrixGuess, InitialGuessModule.jl:11-16).
 * qlb_scf_build        J, K <- P in one fused screened pass (sharded + all-reduced over the communicator's ranks), then
 *                      energies4 = {2 tr(TP), 2 tr(VP), 2 tr(JP), -tr(KP)}: computeEnergyHartreeFock (:25-35), Vnn excluded.
 *                      incremental != 0: J(P) = J(P_prev) + J(P - P_prev) with a full rebuild every 8th build (8f-3).
 * qlb_scf_step         F = T + V + 2J - K; eigen(Symmetric(F), Symmetric(S)) (cuSOLVER Dsygvd); P = Cocc Cocc^T with the
 *                      library's own DMMA GEMM: evaluateSCFStep (:73-84).
 * qlb_scf_get          what = 0 P, 1 F, 2 J, 3 K (N x N), 4 orbital energies (N), 5 orbital coefficients (N x N). */
typedef struct qlb_scf qlb_scf;
int qlb_scf_create(qlb_basis *b, const double *S, const double *T, const double *V, int nocc, qlb_scf **out);
int qlb_scf_destroy(qlb_scf *s);
int qlb_scf_set_density(qlb_scf *s, const double *P);
int qlb_scf_build(qlb_scf *s, double tau, int incremental, double *energies4, qlb_stats *stats);
int qlb_scf_step(qlb_scf *s);
int qlb_scf_get(qlb_scf *s, int what, double *out);
/* C = alpha op(A) op(B) + beta C on DEVICE pointers (column-major) with the library's hand-written FP64 tensor-core
 * kernel (mma.sync m8n8k4 f64 = DMMA): the contraction engine of qlb_scf_step and of the RI path; exported for tests. */
int qlb_dgemm_device(int transA, int transB, int M, int N, int K, double alpha, const double *dA, int lda, const double *dB,
                     int ldb, double beta, double *dC, int ldc, void *stream);
/* The packed-symmetric products of the RI-K contraction (ResolutionIdentityModule.jl:69-78) on DEVICE pointers, exported
 * for tests: dBp = nslab packed lower triangles (column-major, N(N+1)/2 each).  side 0: X[:,:,q] = B^q R (R N x ncol);
 * side 1: out = sum_q X[:,:,q] B^q (R = X, N x N x nslab). */
int qlb_dsymm_packed_device(int side, int N, int ncol, int nslab, const double *dBp, const double *dR, double *dOut, void *stream);

/* ---- knobs ---------------------------------------------------------------------------------------------- */
int qlb_qlog(double x);        /* the integer screening rule's qlog(x) = ceil(8 log2 x), exact table form (tests) */
int qlb_set_profiling(int on); /* time every class kernel with its own event pair (serialises the classes) */
int qlb_fp64_peak_probe(double *tflops); /* DFMA microbenchmark: the measured FP64 roofline denominator */
int qlb_fp64_tensor_peak_probe(double *tflops); /* DMMA (mma.sync m8n8k4 f64) microbenchmark: the RI contractions' denominator */
/* Test hook: injects the schedule feedback a multi-rank run would have all-reduced (counts[2*qc] quartets,
 * counts[2*qc+1] primitive quartets, class_ms[qc] calibration time of class pair qc), so that ONE device can evaluate
 * every piece of the nranks-way deal with qlb_build_jk_device(rank = 0..nranks-1) and check that the pieces sum up. */
int qlb_debug_set_schedule_feedback(qlb_basis *b, int nranks, const double *counts, const double *class_ms);

#ifdef __cplusplus
}
#endif
#endif /* QLB200_H */
